"""Configuration tree with the same key schema and default values as the reference's yacs tree
(reference model/config/defaults.py:14-121), readable as `cfg.SECTION.KEY`, mergeable from the
reference's YAML files unchanged (config/config_csbsr_pspnet.yaml).  yacs itself is not required:
`CfgNode` below implements the subset of its API the entry points use
(merge_from_file / merge_from_list / freeze / defrost / clone / dump)."""
import ast
import copy

import yaml


def _decode(value):
    """Strings are passed through ast.literal_eval when possible (as yacs does), so `LR: 2e-5` becomes a float."""
    if isinstance(value, str):
        try:
            return ast.literal_eval(value)
        except (ValueError, SyntaxError):
            return value
    return value


class CfgNode(dict):
    def __init__(self, init=None):
        super().__init__()
        object.__setattr__(self, "_frozen", False)
        for k, v in (init or {}).items():
            dict.__setitem__(self, k, CfgNode(v) if isinstance(v, dict) else v)

    # attribute access ---------------------------------------------------------------------
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __setitem__(self, key, value):
        if object.__getattribute__(self, "_frozen"):
            raise AttributeError("attempt to modify frozen CfgNode key %r" % (key,))
        dict.__setitem__(self, key, CfgNode(value) if isinstance(value, dict) and not isinstance(value, CfgNode) else value)

    # yacs-like API ------------------------------------------------------------------------
    def _merge(self, other, path=""):
        for k, v in other.items():
            if k not in self:
                raise KeyError("non-existent config key: %s%s" % (path, k))
            if isinstance(self[k], CfgNode):
                if not isinstance(v, dict):
                    raise ValueError("config key %s%s expects a mapping" % (path, k))
                self[k]._merge(v, path + k + ".")
            else:
                self[k] = _decode(v)

    def merge_from_file(self, filename):
        with open(filename) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, pairs):
        assert len(pairs) % 2 == 0
        for key, value in zip(pairs[0::2], pairs[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            if parts[-1] not in node:
                raise KeyError("non-existent config key: %s" % key)
            node[parts[-1]] = _decode(value)

    def _set_frozen(self, flag):
        object.__setattr__(self, "_frozen", flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def is_frozen(self):
        return object.__getattribute__(self, "_frozen")

    def clone(self):
        c = CfgNode(self._plain())
        return c

    def _plain(self):
        return {k: (v._plain() if isinstance(v, CfgNode) else copy.deepcopy(v)) for k, v in self.items()}

    def dump(self, **kw):
        return yaml.safe_dump(self._plain(), **kw)

    def __deepcopy__(self, memo):
        return self.clone()


_CRACK = "datasets/crack_segmentation_dataset/"

_DEFAULTS = {
    "DEVICE": "cuda",
    "MODEL": {
        "SCALE_FACTOR": 4, "DETECTOR_TYPE": "u-net16", "SR": "DBPN", "UP_SAMPLE_METHOD": "deconv",
        "DETECTOR_DBPN_NUM_STAGES": 4, "OPTIMIZER": "Adam", "NUM_CLASSES": 1, "NUM_STAGES": 4,
        "SR_SEG_INV": False, "JOINT_LEARNING": True, "SR_RESIDUAL_LEARNING": True, "KBPN_KERNEL_SFT": True,
        "SR_PIXEL_SHUFFLE": False, "SR_SCRATCH": False, "DSRL_UPSAMPLE": "bilinear", "SUM_LR_ERROR_POS": "HR",
        "ZERO_PAD_KERNEL": False,
    },
    "SOLVER": {
        "MAX_ITER": 300000, "TRAIN_DATASET_RATIO": 0.95,
        "SR_PRETRAIN_ITER": [1, 150001], "SR_SR_MODULE_PRETRAIN_ITER": [1, 50001],
        "SR_KERNEL_MODULE_PRETRAIN_ITER": [50001, 100000], "ONLY_KERNEL_LOSS_FOR_PRETRAIN": False,
        "SEG_PRETRAIN_ITER": [0, 0], "BATCH_SIZE": 8,
        "TASK_LOSS_WEIGHT": 0.5, "INCRESE_TASK_W_ITER": [30000, 170000],
        "SEG_LOSS_FUNC": "Dice", "BOUNDARY_DEC_RATIO": 1.0, "WB_AND_D_WEIGHT": [1, 1], "BCELOSS_WEIGHT": [20, 1],
        "SEG_AUX_LOSS_WEIGHT": 0.4, "SEG_MAIN_LOSS_WEIGHT": 1.0,
        "DSRL_FA_WEIGHT": 0.5, "DSRL_SR_WEIGHT": 0.5, "DSRL_SEG_WEIGHT": 1.0,
        "ORIENTED_WEIGHT_GAUS": 2, "ORIENTED_WEIGHT_ITER": -1,
        "CRACK_ORIENTED_WEIGHT4SR_AMP": 0.0, "CRACK_ORIENTED_WEIGHT4SR_BIAS": 1.0,
        "CRACK_ORIENTED_WEIGHT4SS_AMP": 0.0, "CRACK_ORIENTED_WEIGHT4SS_BIAS": 1.0,
        "SEG_FAIL_ORIENTED_WEIGHT4SR_AMP": 0.0, "SEG_FAIL_ORIENTED_WEIGHT4SR_BIAS": 1.0,
        "SEG_FAIL_ORIENTED_WEIGHT4SS_AMP": 0.0, "SEG_FAIL_ORIENTED_WEIGHT4SS_BIAS": 1.0,
        "INTERM_SSLOSSWEGHT4SR": False,
        "SR_LOSS_FUNC": "L1",
        # NB the reference default really is the 4-element [0.4, 0.4, 0, 2] (a `0,2` typo for 0.2):
        # the kernel-MSE weight is therefore 0 -- reproduced on purpose.
        "SR_LOSS_FUNC_SR_WEIGHT": [0.4, 0.4, 0, 2],
        "LR_LOSS_FUNC": "L1", "ALPHA_MIN": 0.01, "DECREASE_RATIO": 1.0, "SYNC_BATCHNORM": True,
        "NORM_SR_OUTPUT": "all", "LR": 1e-3, "LR_STEPS": [], "SCHEDULER": True, "GAMMA": 0.1,
        "WARMUP_FACTOR": 1.0, "WARMUP_ITERS": 5000, "DOWNSCALE_INTERPOLATION": "bicubic",
    },
    "BLUR": {"FLAG": True, "KERNEL_SIZE": 21, "KERNEL_SIZE_OUTPUT": 21, "ISOTROPIC": False},
    "INPUT": {"IMAGE_SIZE": [448, 448], "MEAN": [0.4741, 0.4937, 0.5048], "STD": [0.1621, 0.1532, 0.1523]},
    "DATASET": {
        "ONLY_IMAGES": False,
        "DATA_AUGMENTATION": [["ConvertFromInts", None], ["RandomMirror", None], ["ToTensor", None],
                              ["RandomVerticalFlip", {"p": 0.3}],
                              ["RandomResizedCrop", {"scale": (1.0, 1.0), "ratio": (1.0, 1.0)}]],
        "TRAIN_IMAGE_DIR": _CRACK + "train/images", "TRAIN_MASK_DIR": _CRACK + "train/masks",
        "TEST_IMAGE_DIR": _CRACK + "test_blured/gt/images", "TEST_MASK_DIR": _CRACK + "test_blured/gt/masks",
        "TEST_BLURED_DIR": _CRACK + "test_blured/", "TEST_BLURED_NAME": "02_40",
    },
    "OUTPUT_DIR": "output/CSSR_SR-SS",
    "SEED": 1121,
    "BASE_NET": "weights/vgg16_reducedfc.pth",
}


def get_cfg_defaults():
    return CfgNode(_DEFAULTS)
