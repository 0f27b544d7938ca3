from .defaults import CfgNode, get_cfg_defaults

cfg = get_cfg_defaults()

__all__ = ["cfg", "CfgNode", "get_cfg_defaults"]
