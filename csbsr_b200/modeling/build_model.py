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
        self.chunk = 16                     # images per pass through KBPN (activation working set; 16 fills whole waves of CTA pairs)
        self.seg_chunk = 32                 # images per pass through the segmentation net
        self._engines = None
        self._packed_version = None

    # -------------------------------------------------------------- weights
    def _param_version(self):
        # tensor versions catch in-place torch updates; the fused Adam / BatchNorm kernels write through raw pointers and do
        # not bump them, so the optimizer's own change counter (kernels.invalidate_packed, called by FusedAdam.step) is part
        # of the key: the eval engines are re-packed after every training step even with frozen BatchNorm buffers
        from .. import kernels as K
        return (K._PACK_STATE["epoch"],) + tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())

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


class BoundaryComboSchedule:
    """The attributes of BoundaryComboLoss the trainer pokes (loss_functions.py:26-41, 76-81; trainer.py:93-96,
    497-508): alpha starts at 1 and drops by 0.01*decrease_ratio per epoch down to alpha_min."""

    def __init__(self, per_epoch, resume_iter=0, alpha_min=0.01, decrease_ratio=1.0):
        self.alpha_min, self.decrease_ratio, self.per_epoch = alpha_min, decrease_ratio, per_epoch
        self.fix_alpha = False
        self.iter = resume_iter % per_epoch
        self.alpha = max(alpha_min, 1.0 - (resume_iter // per_epoch) * 0.01 * decrease_ratio)

    def update_alpha(self):
        if self.iter % self.per_epoch == 0 and self.alpha > self.alpha_min and not self.fix_alpha:
            self.alpha -= 0.01 * self.decrease_ratio
            self.iter = 1
        else:
            self.iter += 1


class JointModelWithLoss(JointModel):
    """JointModelWithLoss(cfg, num_train_ds, resume_iter, sr_transforms).forward(iter, x, sr_targets, segment_targets,
    kernel_targets) -> (segment_loss, sr_loss, segment_preds, sr_preds, kernel_preds) -- reference build_model.py:323-416.

    The returned losses carry an autograd graph whose conv / transposed-conv nodes run forward, dgrad and wgrad on the
    tcgen05 engine (csbsr_b200/autograd.py, modeling/train_graph.py).  Supported: KBPN + PSPNet outside the pre-training
    phases (iteration >= SOLVER.SR_PRETRAIN_ITER[1], all modules trainable), SEG_LOSS_FUNC 'BoundaryCombo', SR_LOSS_FUNC
    'KBPN'.  In eval mode (`.eval()`), forward falls through to JointModel's engines."""

    def __init__(self, cfg, num_train_ds, resume_iter=0, sr_transforms=None):
        super().__init__(cfg)
        if self.seg_model_name not in ("PSPNet", "HRNet_OCR", "PSPNet_BlurSkip"):
            raise NotImplementedError("training graph: DETECTOR_TYPE=%r" % (self.seg_model_name,))
        if self.seg_model_name == "PSPNet_BlurSkip":             # only the BlurSkip branch is trained (build_model.py:352-366)
            for name, p_ in self.named_parameters():
                p_.requires_grad_(name.startswith("segmentation_model.blur_skip."))
        if cfg.SOLVER.SEG_LOSS_FUNC != "BoundaryCombo" or cfg.SOLVER.SR_LOSS_FUNC != "KBPN":
            raise NotImplementedError("training graph: SEG_LOSS_FUNC=%r SR_LOSS_FUNC=%r" %
                                      (cfg.SOLVER.SEG_LOSS_FUNC, cfg.SOLVER.SR_LOSS_FUNC))
        # csrc/losses.cu implements BoundaryComboLoss with pos_weight = loss_weight = [1, 1] (the shipped YAMLs); the
        # reference feeds both keys through (build_model.py:293-299), so any other value must not train silently
        if list(cfg.SOLVER.BCELOSS_WEIGHT) != [1, 1] or list(cfg.SOLVER.WB_AND_D_WEIGHT) != [1, 1]:
            raise NotImplementedError("training graph: BCELOSS_WEIGHT=%r WB_AND_D_WEIGHT=%r (only [1, 1] is built)" %
                                      (list(cfg.SOLVER.BCELOSS_WEIGHT), list(cfg.SOLVER.WB_AND_D_WEIGHT)))
        if not cfg.BLUR.FLAG or cfg.BLUR.ISOTROPIC:
            raise NotImplementedError("training graph: BLUR.FLAG=%r BLUR.ISOTROPIC=%r (anisotropic on-the-fly blur only)" %
                                      (cfg.BLUR.FLAG, cfg.BLUR.ISOTROPIC))
        self.cfg = cfg
        self.main_weight, self.aux_weight = cfg.SOLVER.SEG_MAIN_LOSS_WEIGHT, cfg.SOLVER.SEG_AUX_LOSS_WEIGHT
        self.wf_amp = cfg.SOLVER.SEG_FAIL_ORIENTED_WEIGHT4SS_AMP
        self.oriented_w_iter = cfg.SOLVER.ORIENTED_WEIGHT_ITER
        self.sr_loss_weights = tuple(cfg.SOLVER.SR_LOSS_FUNC_SR_WEIGHT)          # [HR, LR, kernel, -] defaults.py:72
        sr_pre = cfg.SOLVER.SR_PRETRAIN_ITER
        seg_rsm = resume_iter - (sr_pre[1] - 1) if resume_iter > (sr_pre[1] - 1) else 0
        self.ss_loss_fn = BoundaryComboSchedule(num_train_ds // cfg.SOLVER.BATCH_SIZE + 1, seg_rsm,
                                                decrease_ratio=cfg.SOLVER.BOUNDARY_DEC_RATIO)
        self.iter_cnt = True
        self.dropout = True                       # parity harness switches Dropout2d off (masks are random)
        self.freeze_bn = False                    # True: BatchNorm uses its running statistics in train mode too

    def apply_phase(self, iter):
        """requires_grad switches of KBPN._pretrain_check / KBlock._pretrain_check (kbpn.py:118-155, 414-447) as a function
        of the iteration.  Returns True while the ground-truth kernel drives KBPN (SR-module pre-training)."""
        s = self.cfg.SOLVER
        in_sr = s.SR_SR_MODULE_PRETRAIN_ITER[0] <= iter < s.SR_SR_MODULE_PRETRAIN_ITER[1]
        k_lo, k_hi = s.SR_KERNEL_MODULE_PRETRAIN_ITER
        in_k = k_lo <= iter < k_hi
        in_k_early = k_lo <= iter < k_hi - 1          # KBPN re-enables its SR layers one iteration early (kbpn.py:135)
        for name, p in self.sr_model.named_parameters():
            if ".kb.kernel_predictor." in name:
                p.requires_grad_(not in_sr)
            elif ".kb.sr_reconst." in name or ".kb.up_conv1." in name:
                p.requires_grad_(not in_k)
            elif name.startswith("predictor."):
                p.requires_grad_(True)
            else:                                      # feat, UpBlock / DownBlock / SFTlayer of every stage, output_conv
                p.requires_grad_(not in_k_early)
        return in_sr

    def _tensors(self):
        t = dict(self.named_parameters())
        t.update(dict(self.named_buffers()))
        return t

    def forward(self, iter, x, sr_targets=None, segment_targets=None, kernel_targets=None):
        from ..engine import losses as LS
        from . import train_graph as TG
        if not torch.cuda.is_available() or not _lib.lib().csbsr_device_ok():
            raise _lib.CsbsrError("csbsr_b200 needs an sm_100 CUDA device; there is no CPU fallback")
        cfg = self.cfg
        sr_module_pre = self.apply_phase(iter) if self.seg_model_name != "PSPNet_BlurSkip" else False
        # loss = sr_loss inside SR_PRETRAIN_ITER unless SEG_PRETRAIN_ITER overrides it with segment_loss (trainer.py:432-437);
        # the forward graph itself does not depend on SEG_PRETRAIN_ITER
        sr_only = cfg.SOLVER.SR_PRETRAIN_ITER[0] <= iter < cfg.SOLVER.SR_PRETRAIN_ITER[1] and not \
            cfg.SOLVER.SEG_PRETRAIN_ITER[0] <= iter < cfg.SOLVER.SEG_PRETRAIN_ITER[1]
        device = torch.device("cuda", torch.cuda.current_device())
        mv = lambda t: None if t is None else t.to(device=device, dtype=torch.float32)
        x, sr_targets, segment_targets, kernel_targets = mv(x), mv(sr_targets), mv(segment_targets), mv(kernel_targets)
        P = self._tensors()
        was = torch.backends.cudnn.enabled
        torch.backends.cudnn.enabled = False        # glue ops (BN, pooling, resampling) stay on native aten kernels
        try:
            sr, kvec = TG.kbpn_forward(P, x, self.num_stages, self.ksize, self.scale_factor,
                                       gt_kernel=kernel_targets if sr_module_pre else None)
            seg_fwd = TG.hrnet_ocr_forward if self.seg_model_name == "HRNet_OCR" else TG.pspnet_forward
            # while only sr_loss is optimised the segmentation net is evaluated for logging only: no tape needed
            extra = {"kvec": kvec} if self.seg_model_name == "PSPNet_BlurSkip" else {}
            with torch.set_grad_enabled(torch.is_grad_enabled() and not sr_only):
                from .. import glue as G
                normed = G.instance_norm(sr, eps=1e-5)                            # norm_sr, build_model.py:135-137
                # every segmentation-net layer is downstream of `normed`: when its gradient arrives, all of their backward
                # kernels (weight gradients folded into the flat buffer) are enqueued -- the trainer starts the NCCL
                # all-reduce of that part of the gradient here, under the backward of the SR net (engine/trainer.py)
                cb = getattr(self, "on_seg_backward_done", None)
                if cb is not None and normed.requires_grad and self.seg_model_name != "PSPNet_BlurSkip":
                    normed.register_hook(lambda g_, cb=cb: cb())
                drop = None
                if self.dropout and self.training:
                    if getattr(self, "_drop_state", None) is None or self._drop_state.counter.device != device:
                        self._drop_state = G.DropoutState(device, seed=cfg.SEED)
                    drop = self._drop_state
                    drop.begin_step()
                seg, aux = seg_fwd(P, normed, bn_training=self.training and not self.freeze_bn, dropout=drop, **extra)
            sr_loss, kernel_preds = LS.kbpn_loss_train(sr, sr_targets, x, kvec, kernel_targets, self.sr_loss_weights,
                                                       self.ksize, self.scale_factor)
            amp = self.wf_amp if self.oriented_w_iter <= iter else 0.0
            if self.wf_amp != 0 and amp == 0:
                raise NotImplementedError("out_map loss without the w^F weight (iteration < ORIENTED_WEIGHT_ITER)")
            seg_loss = LS.seg_loss_train(seg, aux, segment_targets, self.ss_loss_fn.alpha, amp, self.main_weight,
                                         self.aux_weight)
        finally:
            torch.backends.cudnn.enabled = was
        return seg_loss, sr_loss, seg, sr, kernel_preds
