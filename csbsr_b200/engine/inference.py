"""Evaluation metrics of the reference's inference loop (model/engine/inference.py:25-207) on the device.

`seg_metrics` replaces, for one batch, the threshold sweep (:49-53,111), IoU (:119 ->
model/utils/estimate_metrics.py:72-84) and calc_distance_metrics (:121 -> :293-336): it never materialises
the (B,99,H,W) binarised tensor and returns integer IoU counts plus fp64 HD / MSD values that are
bit-identical to the reference's.  The fp64 ratio/mean arithmetic on the tiny (B,99) results is the same
numpy expression the reference uses (:171-173)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib

NUM_THRESHOLDS = 99
# torch.Tensor([i*0.01 for i in range(1,100)]) -> float32 (inference.py:49-51)
THRESHOLDS = np.array([i * 0.01 for i in range(1, 100)], dtype=np.float64).astype(np.float32)
HD_PERCENTILE = 50        # the shipped reference evaluates "HD95" with percentile = 50 (inference.py:302)

_thr_cache = {}
_ws_cache = {}


def _thresholds(device):
    t = _thr_cache.get(device)
    if t is None:
        t = torch.from_numpy(THRESHOLDS.copy()).to(device)
        _thr_cache[device] = t
    return t


def seg_metrics(segment_preds, masks, with_hd=True, percent=HD_PERCENTILE, to_host=True):
    """segment_preds, masks: fp32 [B,1,H,W] (any device; moved to the current CUDA device).
    Returns dict(inter, union [B,99] int64; iou [B,99] float64; hd, msd [B,99] float64 or None)."""
    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    p = segment_preds.to(device=dev, dtype=torch.float32).contiguous()
    m = masks.to(device=dev, dtype=torch.float32).contiguous()
    B, c, H, W = p.shape
    assert c == 1 and m.shape == p.shape
    inter = torch.empty((B, NUM_THRESHOLDS), dtype=torch.int64, device=dev)
    union = torch.empty_like(inter)
    hd = torch.empty((B, NUM_THRESHOLDS), dtype=torch.float64, device=dev) if with_hd else None
    msd = torch.empty_like(hd) if with_hd else None
    need = L.csbsr_metrics_workspace_bytes(B, H, W, int(with_hd))
    key = (dev, B, H, W, bool(with_hd))
    ws = _ws_cache.get(key)
    if ws is None:
        _ws_cache.clear()
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    rc = L.csbsr_seg_metrics(p.data_ptr(), m.data_ptr(), _thresholds(dev).data_ptr(), B, H, W, inter.data_ptr(),
                             union.data_ptr(), hd.data_ptr() if with_hd else None, msd.data_ptr() if with_hd else None,
                             C.c_double(float(percent)), ws.data_ptr(), need, _lib.stream_ptr())
    _lib.check(rc, "csbsr_seg_metrics")
    _lib.count_launch("csbsr_seg_metrics", with_hd)
    out = {"inter": inter, "union": union, "hd": hd, "msd": msd}
    if to_host:
        out = {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}
        out["iou"] = (out["inter"] + 1e-5) / (out["union"] + 1e-5)      # estimate_metrics.py:75,84
    return out


class IoU:
    """Same call convention as the reference's IoU for the swept case: __call__(prob, target) -> (B,99) ndarray."""

    def __init__(self, th=0.5):
        self.name = "IoU"
        self.th = 0.5

    def __call__(self, segment_preds, target):
        return seg_metrics(segment_preds, target, with_hd=False)["iou"]


def calc_distance_metrics(segment_preds, gts, num_hd_outliner=0, num_msd_outliner=0, percent=HD_PERCENTILE):
    """Same return convention as the reference (inference.py:293-336): (hd[B,99], msd[B,99], n_hd_outliers,
    n_msd_outliers); an "outlier" is a (image, threshold) pair where exactly one of gt / prediction is empty."""
    r = seg_metrics(segment_preds, gts, with_hd=True, percent=percent)
    g = gts.to(dtype=torch.float32)
    tgt_iou = (g > 0.5).flatten(1).sum(1).cpu().numpy().astype(np.int64)[:, None]
    gt_empty = (~(g != 0).flatten(1).any(1)).cpu().numpy()[:, None]
    pred_empty = (r["union"] - tgt_iou + r["inter"]) == 0
    one_empty = gt_empty ^ pred_empty
    n = int(one_empty.sum())
    return r["hd"], r["msd"], num_hd_outliner + n, num_msd_outliner + n


# ------------------------------------------------------------------ evaluation loop (host glue around the kernels)
def psnr_ssim(pred, target):
    """PSNR()(pred, target) and SSIM()(pred, target) of the reference (model/utils/estimate_metrics.py:89-100, 134-191) for
    [B,C,H,W] tensors in [0,1]: one csbsr_psnr_ssim launch -> (psnr float64 [B], ssim float64 [B]) numpy arrays."""
    import ctypes as C
    dev = torch.device("cuda", torch.cuda.current_device())
    a = pred.to(device=dev, dtype=torch.float32).contiguous()
    b = target.to(device=dev, dtype=torch.float32).contiguous()
    n, c, h, w = a.shape
    out = torch.empty(2, n, dtype=torch.float64, device=dev)
    L = _lib.lib()
    nb = L.csbsr_psnr_ssim_workspace_bytes(n)
    ws = torch.empty(max(int(nb), 8), dtype=torch.uint8, device=dev)
    _lib.check(L.csbsr_psnr_ssim(a.data_ptr(), b.data_ptr(), n, c, h, w, out[0].data_ptr(), out[1].data_ptr(), ws.data_ptr(),
                                 C.c_size_t(nb), _lib.stream_ptr()), "csbsr_psnr_ssim")
    _lib.count_launch("csbsr_psnr_ssim")
    res = out.cpu().numpy()
    return res[0], res[1]


def psnr_per_image(pred, target):
    return psnr_ssim(pred, target)[0]


class PSNR:
    """Drop-in for estimate_metrics.PSNR (returns a numpy array per image)."""
    name = "PSNR"

    def __call__(self, img1, img2):
        return psnr_ssim(img1, img2)[0].astype(np.float32)


class SSIM:
    """Drop-in for estimate_metrics.SSIM(window_size=11, size_average=False)."""

    def __call__(self, img1, img2):
        return psnr_ssim(img1, img2)[1].astype(np.float32)

    forward = __call__


def inference_for_ss(model, loader, test_surface_distance=True, percent=HD_PERCENTILE, output_dir=None, log=print):
    """Counterpart of the reference's inference_for_ss (model/engine/inference.py:25-207) for loaders that yield
    (lr_imgs[B,3,h,w], sr_targets[B,3,H,W], masks[B,1,H,W], kernel_targets[B,1,k,k], fnames): runs the model, the AIU
    sweep and (optionally) the HD/MSD sweep per batch, prints the running means and writes iou_log.csv.
    Under torch.distributed every rank evaluates its own loader shard; results are gathered on all ranks."""
    import os
    from . import distributed as D
    rows, fnames, psnrs, kpsnrs, ssims = [], [], [], [], []
    n_hd_out = 0
    for it, (imgs, sr_targets, masks, kernel_targets, names) in enumerate(loader, 1):
        sr, seg, kp = model(imgs, torch.zeros(imgs.shape[0], 1, model.blur_ksize, model.blur_ksize), sr_targets=sr_targets)
        ps, ss = psnr_ssim(sr, sr_targets)                              # inference.py:96-97
        psnrs.append(ps)
        ssims.append(ss)
        kpsnrs.append(psnr_per_image(torch.clamp(kp, min=0.0, max=1.0), kernel_targets))   # clip, inference.py:97-100
        r = seg_metrics(seg, masks, with_hd=test_surface_distance, percent=percent, to_host=False)
        hd = r["hd"] if test_surface_distance else torch.zeros_like(r["inter"], dtype=torch.float64)
        msd = r["msd"] if test_surface_distance else torch.zeros_like(hd)
        rows.append(D.pack_metrics(r["inter"], r["union"], hd, msd))
        fnames += list(names)
        if it % 10 == 0:
            log("batch %d done" % it)
    # one exchange: the [B, 4*99] metric rows plus three per-image columns (PSNR, SSIM, kernel PSNR), so that every
    # reported mean is over ALL images, not over this rank's shard
    per_img = torch.from_numpy(np.stack([np.concatenate(psnrs), np.concatenate(ssims), np.concatenate(kpsnrs)], 1)
                               .astype(np.float64)).to(rows[0].device)
    packed = D.gather_rows(torch.cat([torch.cat(rows, 0), per_img], 1))
    inter, union, hd, msd = D.unpack_metrics(packed[:, :-3])
    extra = packed[:, -3:].cpu().numpy()
    fnames = D.gather_names(fnames)
    iou = (inter + 1e-5) / (union + 1e-5)
    out = {"AIU": float(np.mean(iou)), "IoU_max": float(np.max(np.mean(iou, axis=0))), "iou": iou,
           "PSNR": float(np.mean(extra[:, 0])), "SSIM": float(np.mean(extra[:, 1])),
           "PSNR_kernel": float(np.mean(extra[:, 2]))}
    if test_surface_distance:
        out.update({"AHD": float(np.mean(hd)), "HD_min": float(np.min(np.mean(hd, axis=0))), "AMSD": float(np.mean(msd)),
                    "hd": hd, "msd": msd})
    log("estimation finish!!  PSNR_mean:%.4f  SSIM_mean:%.4f PSNR(Kernel)_mean:%.4f AIU_mean:%.4f"
        % (out["PSNR"], out["SSIM"], out["PSNR_kernel"], out["AIU"])
        + ("  HD%d_mean:%.4f MSD_mean:%.4f" % (percent, out["AHD"], out["AMSD"]) if test_surface_distance else ""))
    rank, _ = D.world()
    if output_dir and rank == 0:
        os.makedirs(output_dir, exist_ok=True)
        with open(os.path.join(output_dir, "iou_log.csv"), "w") as f:       # save_iou_log, inference.py:287-291
            f.write("," + ",".join("%g" % (i * 0.01) for i in range(1, 100)) + "\n")
            for i in range(iou.shape[0]):
                name = fnames[i] if i < len(fnames) else "image_%d" % i
                f.write(name + "," + ",".join(repr(float(v)) for v in iou[i]) + "\n")
    return out
