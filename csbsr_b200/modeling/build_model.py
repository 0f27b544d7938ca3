"""Drop-in model builder: same class names, constructor arguments, forward signatures and state_dict keys
as the reference's model/modeling/build_model.py (JointModel :441-500), with the forward executed by the
sm_100a engines in this package.  There is no CPU / eager fallback: forward raises when the CUDA
library is missing or no sm_100 device is visible."""
import torch
from torch import nn

from .. import _lib
from .. import kernels as K
from .kbpn import KBPNEngine
from .hrnet_ocr import HRNetOCREngine
from .params import ParamTree, hrnet_ocr_param_shapes, kbpn_param_shapes, pspnet_param_shapes
from .pspnet import PSPNetEngine


class JointModel(nn.Module):
    """JointModel(cfg).forward(x, damy_kernel, sr_targets=None) -> (sr_preds[B,3,4h,4w], segment_preds[B,1,4h,4w],
    kernel_preds[B,1,k,k]) -- reference build_model.py:466-496 (KBPN + PSPNet path, NORM_SR_OUTPUT 'instance')."""

    def __init__(self, cfg):
        super().__init__()
        if cfg.MODEL.SR != "KBPN":
            raise NotImplementedError(cfg.MODEL.SR)
        if cfg.MODEL.DETECTOR_TYPE not in ("PSPNet", "PSPNet_BlurSkip", "HRNet_OCR"):
            raise NotImplementedError(cfg.MODEL.DETECTOR_TYPE)
        if cfg.MODEL.SCALE_FACTOR != 4:
            raise NotImplementedError("SCALE_FACTOR=%r" % (cfg.MODEL.SCALE_FACTOR,))
        if cfg.SOLVER.NORM_SR_OUTPUT != "instance":
            raise NotImplementedError("NORM_SR_OUTPUT=%r" % (cfg.SOLVER.NORM_SR_OUTPUT,))
        self.scale_factor = cfg.MODEL.SCALE_FACTOR
        self.ksize = cfg.BLUR.KERNEL_SIZE_OUTPUT
        self.blur_ksize = cfg.BLUR.KERNEL_SIZE
        self.num_stages = cfg.MODEL.NUM_STAGES
        self.seg_model_name = cfg.MODEL.DETECTOR_TYPE
        self.norm_method = cfg.SOLVER.NORM_SR_OUTPUT
        blur_dim = cfg.BLUR.KERNEL_SIZE_OUTPUT ** 2 if self.seg_model_name == "PSPNet_BlurSkip" else None
        if self.seg_model_name == "HRNet_OCR":
            self.segmentation_model = ParamTree(hrnet_ocr_param_shapes(cfg.MODEL.NUM_CLASSES))
        else:
            self.segmentation_model = ParamTree(pspnet_param_shapes(cfg.MODEL.NUM_CLASSES, blur_dim=blur_dim))
        self.sr_model = ParamTree(kbpn_param_shapes(self.num_stages, 128, self.blur_ksize, self.ksize))
        self.chunk = 8                      # images per pass through KBPN (activation working set)
        self.seg_chunk = 32                 # images per pass through the segmentation net
        self._engines = None
        self._packed_version = None

    # -------------------------------------------------------------- weights
    def _param_version(self):
        return tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._packed_version = None
        return out

    def _ensure_engines(self, device):
        ver = (str(device), self._param_version())
        if self._engines is None or self._packed_version != ver:
            sd = self.state_dict()
            sr = KBPNEngine(self.num_stages, 128, self.blur_ksize, self.ksize, self.scale_factor, device).load(sd)
            ss = (HRNetOCREngine(device=device) if self.seg_model_name == "HRNet_OCR" else PSPNetEngine(device=device)).load(sd)
            self._engines = (sr, ss)
            self._packed_version = ver
        return self._engines

    # -------------------------------------------------------------- forward
    @torch.no_grad()
    def forward(self, x, damy_kernel=None, sr_targets=None, return_aux=False):
        if not torch.cuda.is_available() or not _lib.lib().csbsr_device_ok():
            raise _lib.CsbsrError("csbsr_b200 needs an sm_100 CUDA device; there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
        x = x.to(device=device, dtype=torch.float32)          # MetaSRModel.mount_cuda, build_model.py:118-123
        sr_eng, ss_eng = self._ensure_engines(device)
        B = x.shape[0]
        srs, segs, kps, auxs, kvecs, stats = [], [], [], [], [], []
        # KBPN in chunks of `chunk` images (its 448^2 x 512-channel activations dominate the working set) ...
        for i in range(0, B, self.chunk):
            xc = x[i:i + self.chunk].contiguous()
            b = xc.shape[0]
            sr, kvec = sr_eng.forward(xc)
            mean = torch.empty(b * 3, dtype=torch.float32, device=device)
            rstd = torch.empty(b * 3, dtype=torch.float32, device=device)
            K.clip_instnorm_stats(sr, mean, rstd, do_clip=True)            # clip_sr :143-146 + norm_sr stats :135-137
            kp = torch.empty_like(kvec)
            K.vec_normalize(kvec, kp)                                       # :491-494
            srs.append(sr); kvecs.append(kvec); stats.append((mean, rstd))
            kps.append(kp.view(b, 1, self.ksize, self.ksize))
        sr_all = srs[0] if len(srs) == 1 else torch.cat(srs, 0)
        kvec_all = kvecs[0] if len(kvecs) == 1 else torch.cat(kvecs, 0)
        mean_all = torch.cat([m for m, _ in stats]); rstd_all = torch.cat([r for _, r in stats])
        # ... the segmentation net in larger chunks: its 56^2 layers need more images to fill 148 SMs
        blur = self.seg_model_name == "PSPNet_BlurSkip"
        for i in range(0, B, self.seg_chunk):
            j = min(B, i + self.seg_chunk)
            # PSPNet_BlurSkip also receives the (spatially constant) kernel map of KBPN (build_model.py:482-483, 498-500)
            seg, aux = ss_eng.forward(sr_all[i:j], mean_all[3 * i:3 * j].contiguous(), rstd_all[3 * i:3 * j].contiguous(),
                                      kvec=kvec_all[i:j].contiguous() if blur else None)
            segs.append(seg); auxs.append(aux)
        srs = [sr_all]
        cat = (lambda l: l[0] if len(l) == 1 else torch.cat(l, 0))
        if return_aux:
            return cat(srs), cat(segs), cat(kps), cat(auxs)
        return cat(srs), cat(segs), cat(kps)
