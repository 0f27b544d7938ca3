"""KBPN blind-SR forward on the tcgen05 conv engine (eval path, iter = -1).

Mirrors KBPN.forward and its blocks (reference model/modeling/kbpn.py:84-116, :172-189 stage, :382-412
KBlock, :464-469 UpBlock, :484-489 DownBlock, :511-518 SFTlayer, :562-578 kernel predictor) with these
B200-first restructurings (all exact in real arithmetic):
  * torch.cat -> every producer writes straight into its channel slice of concat_h / concat_l;
  * additions/subtractions between blocks (h1+h0, l0-x, h+e_h, l1+l0, f*scale+shift) are conv epilogues;
  * the 8x8 stride-4 transposed convs run as 16 output phases of a 2x2 conv (no zero-stuffing);
  * the blur-kernel map is spatially constant (kbpn.py:404 `vec.expand`), so it is carried as a per-sample
    vector; its passage through fe_kernel (:572-573) and through the 441 conditioning channels of SFT conv0
    (:513-516) is evaluated on a 5x5 / 3x3 "border-class" image and folded into a per-sample, per-class bias
    of the consuming conv (zero padding makes the response differ only within 2 / 1 pixels of the border).
"""
import torch

from .. import kernels as K
from ..kernels import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, Fmap, F32Map


class _Workspace:
    """Named device buffers, allocated once per shape and reused across calls."""

    def __init__(self, device):
        self.device = device
        self.buf = {}

    def fmap(self, name, n, h, w, c, zero=False):
        key = (name, n, h, w, c)
        t = self.buf.get(key)
        if t is None:
            t = torch.zeros((n, h, w, c), dtype=torch.bfloat16, device=self.device)
            self.buf[key] = t
        elif zero:
            t.zero_()
        return Fmap(t)

    def f32(self, name, *shape):
        key = (name,) + tuple(shape)
        t = self.buf.get(key)
        if t is None:
            t = torch.zeros(shape, dtype=torch.float32, device=self.device)
            self.buf[key] = t
        return t


class KBPNEngine:
    def __init__(self, num_stages=4, md_ch=128, k_est=7, k_out=21, scale=4, device="cuda"):
        assert scale == 4, "only the x4 (8/4/2) configuration is built (kbpn.py:23-26)"
        self.S, self.C, self.ke, self.ko, self.scale = num_stages, md_ch, k_est, k_out, scale
        self.device = device
        self.ws = _Workspace(device)
        self.p = None
        self.debug = None                 # set to a dict to capture per-stage intermediates (tests only)
        # kernel predictor as two fused chain kernels (default) or one conv launch per layer (CSBSR_KPRED_FUSED=0, for A/B runs)
        import os
        self.fused_kpred = os.environ.get("CSBSR_KPRED_FUSED", "1") != "0"

    # ------------------------------------------------------------------ weight packing
    def load(self, sd, prefix="sr_model."):
        """Pack all weights from a state_dict (tensors on any device)."""
        dev = "cpu"                       # pack on the host; K.to_device ships the packed tensors (no device kernels at load time)
        g = lambda k: sd[prefix + k].detach().to(dev, torch.float32)
        slope = lambda k: float(sd[prefix + k].detach().float().reshape(-1)[0])
        C, kc, cond = self.C, self.ke * self.ke, self.ko * self.ko
        cond_pad = K.round_up(cond, 64)
        P = {}

        def as1x1(w):  # [Cout, 3, R, S] -> [Cout, R*S*3, 1, 1] matching csbsr_patchify's channel order
            return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1, 1, 1)

        P["feat0"] = K.pack_conv(as1x1(g("feat.0.weight")), g("feat.0.bias"))
        for i in (2, 4, 6):
            P["feat%d" % i] = K.pack_conv(g("feat.%d.weight" % i), g("feat.%d.bias" % i), padding=1)
        for i in range(3):
            P["pred%d" % i] = (K.pack_conv(g("predictor.feat_ext.%d.layer.weight" % i), padding=1, cout_pad=128 if i < 2 else 64),
                               slope("predictor.feat_ext.%d.act.weight" % i))
        for s in range(self.S):
            sp = "back_projection_stages.%d." % s
            st = {}
            st["up.conv"] = (K.pack_conv(g(sp + "up.conv.layer.weight"), g(sp + "up.conv.layer.bias")), slope(sp + "up.conv.act.weight"))
            st["up.d1"] = (K.pack_deconv8s4(g(sp + "up.up_conv1.layer.weight")), slope(sp + "up.up_conv1.act.weight"))
            st["up.c2"] = (K.pack_conv(g(sp + "up.up_conv2.layer.weight"), stride=4, padding=2), slope(sp + "up.up_conv2.act.weight"))
            st["up.d3"] = (K.pack_deconv8s4(g(sp + "up.up_conv3.layer.weight")), slope(sp + "up.up_conv3.act.weight"))
            # sr_reconst of stage s reads concat_h[0:(s+1)C]; by linearity its response is a sum over the C-channel
            # slices, so slice j is read ONCE more after its KBlock update for all later consumers (see forward)
            w_sr = g(sp + "kb.sr_reconst.layer.weight")                  # [3, (s+1)C, 3, 3]
            st["kb.sr_own"] = K.pack_tapexp3x3(w_sr[:, s * C:(s + 1) * C].contiguous())
            kp = sp + "kb.kernel_predictor."
            st["sr0"] = K.pack_conv(as1x1(g(kp + "fe_SR.0.layer.weight")), cout_pad=64, cin_pad=32)   # K = 27 -> 32
            st["sr1"] = K.pack_conv(g(kp + "fe_SR.1.layer.weight"), cout_pad=64)
            for i in (2, 3, 4):        # 32-channel layers: one 32-wide K chunk per tap (64B swizzle rows) halves the MMAs
                st["sr%d" % i] = K.pack_conv(g(kp + "fe_SR.%d.layer.weight" % i), padding=1, cout_pad=64, cin_pad=32)
            st["fk0"] = K.pack_conv(g(kp + "fe_kernel.0.layer.weight"), padding=1, cout_pad=64, cin_pad=cond_pad)
            st["fk1"] = K.pack_conv(g(kp + "fe_kernel.1.layer.weight"), padding=1, cout_pad=64)
            wcat = g(kp + "fe_cat.0.layer.weight")                       # [32, 98, 1, 1]
            st["cat0_sr"] = K.pack_conv(wcat[:, :kc].contiguous(), cout_pad=64)
            st["cat0_k"] = K.pack_conv(wcat[:, kc:].contiguous(), cout_pad=64)
            st["cat1"] = K.pack_conv(g(kp + "fe_cat.1.layer.weight"), padding=1, cout_pad=64, cin_pad=32)
            st["cat2"] = K.pack_conv(g(kp + "fe_cat.2.layer.weight"), padding=1, cout_pad=64, cin_pad=32)
            # the same layers for the fused chain kernels (csrc/kpred_chain.cu), packed on the host into their smem layout
            raw = lambda k: sd[prefix + k].detach().float().cpu()
            st["chain_sr"] = K.pack_chain([(as1x1(raw(kp + "fe_SR.0.layer.weight")), 32, 64),
                                           (raw(kp + "fe_SR.1.layer.weight"), 64, 32),
                                           (raw(kp + "fe_SR.2.layer.weight"), 32, 32),
                                           (raw(kp + "fe_SR.3.layer.weight"), 32, 32),
                                           (raw(kp + "fe_SR.4.layer.weight"), 32, 64)])
            st["chain_cat"] = K.pack_chain([(raw(kp + "fe_cat.0.layer.weight")[:, :kc].contiguous(), 64, 32),
                                            (raw(kp + "fe_cat.1.layer.weight"), 32, 32),
                                            (raw(kp + "fe_cat.2.layer.weight"), 32, 64)])
            # KBlock.up_conv1: ConvTranspose2d(3 -> C, 8, 4, 2) evaluated on the 3x3-patchified LR error:
            # per output phase a 1x1 conv over K = (a*3+b)*3+c with w[c, co, rho_h+6-4a, rho_w+6-4b]
            w = g(sp + "kb.up_conv1.layer.weight")                       # [3, C, 8, 8]
            wp = torch.zeros((16, C, 32), dtype=torch.float32, device=dev)              # K = 27 -> one 32-wide chunk
            for rh in range(4):
                for rw in range(4):
                    for a in range(3):
                        r = rh + 6 - 4 * a
                        if not 0 <= r < 8:
                            continue
                        for b in range(3):
                            q = rw + 6 - 4 * b
                            if not 0 <= q < 8:
                                continue
                            wp[rh * 4 + rw, :, (a * 3 + b) * 3:(a * 3 + b) * 3 + 3] = w[:, :, r, q].t()
            st["kb.d1"] = (K.PackedConv(wp.to(torch.bfloat16).contiguous(), [(0, 0, i) for i in range(16)], 16, 1, 1, 4,
                                        [i // 4 for i in range(16)], [i % 4 for i in range(16)], C, macs_per_pixel=4 * 3 * C),
                           slope(sp + "kb.up_conv1.act.weight"))
            if s < self.S - 1:
                fc = (s + 1) * C
                st["dn.conv"] = (K.pack_conv(g(sp + "down.conv.layer.weight"), g(sp + "down.conv.layer.bias")), slope(sp + "down.conv.act.weight"))
                st["dn.c1"] = (K.pack_conv(g(sp + "down.down_conv1.layer.weight"), stride=4, padding=2), slope(sp + "down.down_conv1.act.weight"))
                st["dn.d2"] = (K.pack_deconv8s4(g(sp + "down.down_conv2.layer.weight")), slope(sp + "down.down_conv2.act.weight"))
                st["dn.c3"] = (K.pack_conv(g(sp + "down.down_conv3.layer.weight"), stride=4, padding=2), slope(sp + "down.down_conv3.act.weight"))
                cc_pad = K.round_up(fc + cond, 64)
                for br in ("scale", "shift"):
                    w0, b0 = g(sp + "sft.SFT_%s_conv0.weight" % br), g(sp + "sft.SFT_%s_conv0.bias" % br)
                    st["sft.%s.0f" % br] = K.pack_conv(w0[:, :fc].contiguous(), padding=1, cout_pad=cc_pad)
                    st["sft.%s.0k" % br] = K.pack_conv(w0[:, fc:].contiguous(), b0, padding=1, cout_pad=cc_pad, cin_pad=cond_pad)
                    st["sft.%s.1" % br] = K.pack_conv(g(sp + "sft.SFT_%s_conv1.weight" % br), g(sp + "sft.SFT_%s_conv1.bias" % br),
                                                      padding=1, cin_pad=cc_pad)
            P[s] = st
        w_out = g("output_conv.layer.weight")                             # [3, S*C, 3, 3]
        for j in range(self.S):
            # consumers of the final slice j: sr_reconst of stages j+1..S-1 (3 channels each), then output_conv
            ws = [g("back_projection_stages.%d.kb.sr_reconst.layer.weight" % c)[:, j * C:(j + 1) * C] for c in range(j + 1, self.S)]
            ws.append(w_out[:, j * C:(j + 1) * C])
            P[j]["kb.sr_later"] = K.pack_tapexp3x3(torch.cat(ws, 0).contiguous())
        self.p = K.to_device(P, self.device)
        return self

    # ------------------------------------------------------------------ forward
    def forward(self, x):
        """x: fp32 [B,3,h,w] on the device -> (sr fp32 [B,3,4h,4w], kernel vector fp32 [B, ko*ko])."""
        assert self.p is not None, "KBPNEngine.load() must be called first"
        P, ws, C = self.p, self.ws, self.C
        B, _, h, w = x.shape
        H, W = h * self.scale, w * self.scale
        kc, cond = self.ke * self.ke, self.ko * self.ko
        cond_pad = K.round_up(cond, 64)
        x = x.contiguous()

        # ---- VGG head (kbpn.py:42-44, :87)
        xp = K.patchify(x, ws.fmap("xp", B, h, w, 64), 3, 3, 1, 1)
        f = K.conv(xp, P["feat0"], ws.fmap("f64a", B, h, w, 64), act=ACT_RELU)
        f = K.conv(f, P["feat2"], ws.fmap("f64b", B, h, w, 64), act=ACT_RELU)
        f = K.conv(f, P["feat4"], ws.fmap("f128a", B, h, w, 128), act=ACT_RELU)
        init_f = K.conv(f, P["feat6"], ws.fmap("init_f", B, h, w, 128), act=ACT_RELU)

        # ---- initial kernel prediction (kbpn.py:320-341)
        z = K.conv(init_f, P["pred0"][0], ws.fmap("f128a", B, h, w, 128), act=ACT_LEAKY, slope=P["pred0"][1])
        z = K.conv(z, P["pred1"][0], ws.fmap("f128b", B, h, w, 128), act=ACT_LEAKY, slope=P["pred1"][1])
        z = K.conv(z, P["pred2"][0], ws.fmap("f64a", B, h, w, 64), act=ACT_LEAKY, slope=P["pred2"][1])
        v49 = K.gap(z, ws.f32("v49", B, kc), kc)
        kvec = K.kernel_update(v49, None, ws.f32("kvec_a", B, cond), self.ke, self.ko, True)
        if self.debug is not None:
            self.debug["init_f"] = init_f.to_nchw_f32()
            self.debug["init_kernel"] = kvec.clone()

        concat_h = ws.fmap("concat_h", B, H, W, self.S * C)
        concat_l = ws.fmap("concat_l", B, h, w, max(1, self.S - 1) * C)
        t0 = ws.fmap("hr_t0", B, H, W, C)
        # sr_acc[:, 3c:3c+3] collects, slice by slice, sr_reconst of stage c+1 (c < S-1) and output_conv (+ the bicubic
        # residual of kbpn.py:113, c = S-1): fp32 planar accumulators
        sr_acc = ws.f32("sr_acc", B, 3 * self.S, H, W)
        up_tail = K.PlanarWin(sr_acc, 0, 3 * self.S)
        sr_acc.zero_()
        K.bicubic_upsample(x, ws.f32("up_res", B, 3, H, W), self.scale)
        sr_acc[:, 3 * (self.S - 1):] = ws.f32("up_res", B, 3, H, W)
        sr = torch.empty((B, 3, H, W), dtype=torch.float32, device=x.device)
        low = init_f
        for s in range(self.S):
            st = P[s]
            hs = concat_h.window(s * C, C)
            # ---- UpBlock (kbpn.py:464-469): h = deconv(l0 - x) + h0 written into its concat_h slice
            xl = K.conv(low, st["up.conv"][0], ws.fmap("lr_x", B, h, w, C), act=ACT_LEAKY, slope=st["up.conv"][1])
            h0 = K.conv(xl, st["up.d1"][0], t0, act=ACT_LEAKY, slope=st["up.d1"][1])
            d = K.conv(h0, st["up.c2"][0], ws.fmap("lr_d", B, h, w, C), act=ACT_LEAKY, slope=st["up.c2"][1], r1=xl, r1_sign=-1.0)
            K.conv(d, st["up.d3"][0], hs, act=ACT_LEAKY, slope=st["up.d3"][1], r1=h0)
            # ---- KBlock (kbpn.py:382-412)
            pre = concat_h.window(0, (s + 1) * C)
            # sr_t = sr_reconst(cat(slices)) = own-slice conv + contributions of the earlier (final) slices in sr_acc
            sr_t = self._conv3x3_few(hs, st["kb.sr_own"], ws.f32("sr_t", B, 3, H, W),
                                     K.PlanarWin(sr_acc, 3 * (s - 1), 3) if s > 0 else None)
            kvec = self._kernel_predictor(st, sr_t, kvec, B, H, W, s)
            if self.debug is not None:
                self.debug["sr_t%d" % s] = sr_t.clone()
                self.debug["kvec%d" % s] = kvec.clone()
            err = K.blur_per_sample(sr_t, kvec, x, ws.f32("err", B, 3, h, w), self.ko, self.scale)
            ep = K.patchify(err, ws.fmap("xp", B, h, w, 64), 3, 3, 1, 1)
            K.conv(ep, st["kb.d1"][0], hs, act=ACT_LEAKY, slope=st["kb.d1"][1], r1=hs)      # h + e_h, in place
            # the slice is final now: add its response to every later consumer (channels [3s, 3S) of sr_acc:
            # sr_reconst of stages s+1.., then output_conv), reading the 448^2 x C slice only once
            later = K.PlanarWin(sr_acc, 3 * s, 3 * (self.S - s))
            if s == self.S - 1:
                self._conv3x3_few(hs, st["kb.sr_later"], sr, later)        # output_conv + bicubic residual -> sr
                break
            self._conv3x3_few(hs, st["kb.sr_later"], later, later if s > 0 else up_tail)
            # ---- DownBlock (kbpn.py:484-489) on concat_h[0:(s+1)C], result into its concat_l slice
            xh = K.conv(pre, st["dn.conv"][0], t0, act=ACT_LEAKY, slope=st["dn.conv"][1])
            l0 = K.conv(xh, st["dn.c1"][0], ws.fmap("lr_x", B, h, w, C), act=ACT_LEAKY, slope=st["dn.c1"][1])
            K.conv(l0, st["dn.d2"][0], t0, act=ACT_LEAKY, slope=st["dn.d2"][1], r1=t0, r1_sign=-1.0)   # h0 - x, in place
            K.conv(t0, st["dn.c3"][0], concat_l.window(s * C, C), act=ACT_LEAKY, slope=st["dn.c3"][1], r1=l0)
            # ---- SFT layer (kbpn.py:511-518)
            low = self._sft(st, concat_l.window(0, (s + 1) * C), kvec, B, h, w, s)
        return sr, kvec.clone()

    def _conv3x3_few(self, x, pc, out, r32):
        """3x3 conv to <= 12 channels as tap expansion: one 1x1 GEMM with N = 9*co, then the shifted gather (fp32 planes)."""
        z = K.conv(x, pc, self.ws.fmap("hr_z%d" % pc.cout_pad, x.n, x.h, x.w, pc.cout_pad))
        return K.tap_gather3x3(z, pc.tap_co, out, r32)

    def _kernel_predictor(self, st, sr_t, kvec, B, H, W, s):
        """KernelPredictorLikeIKC.forward + KBlock renormalisation (kbpn.py:562-578, :391-392)."""
        ws, kc, cond = self.ws, self.ke * self.ke, self.ko * self.ko
        cond_pad = K.round_up(cond, 64)
        # constant branch: fe_kernel on the 5x5 border-class image, folded into fe_cat.0's bias
        small = K.broadcast_vec(kvec, ws.fmap("k5", B, 5, 5, cond_pad))
        a = K.conv(small, st["fk0"], ws.fmap("k5a", B, 5, 5, 64), act=ACT_LEAKY, slope=0.01)
        a = K.conv(a, st["fk1"], ws.fmap("k5b", B, 5, 5, 64), act=ACT_LEAKY, slope=0.01)
        cb = K.conv(a, st["cat0_k"], F32Map(ws.f32("k5bias", B, 5, 5, 64)))
        if self.fused_kpred:
            # image branch + fe_cat + GAP as two fused chains: the 32..64-channel 448^2 intermediates never leave the SM
            a = K.kpred_sr_chain(sr_t, st["chain_sr"], ws.fmap("hr_a64", B, H, W, 64), slope=0.01)
            nws = K._lib.lib().csbsr_kpred_workspace_bytes(B, H, W)
            d49 = K.kpred_cat_chain(a, st["chain_cat"], cb.t, ws.f32("v49", B, kc), ws.f32("kpred_ws", nws // 4), slope=0.01)
            if self.debug is not None:
                self.debug["delta49_%d" % s] = d49.clone()
            out = ws.f32("kvec_b" if (s % 2 == 0) else "kvec_a", B, cond)
            return K.kernel_update(d49, kvec, out, self.ke, self.ko, True)
        # image branch (fe_SR) at HR, one conv launch per layer
        sp = K.patchify(sr_t, ws.fmap("hr_p64", B, H, W, 64), 3, 3, 1, 1)
        a = K.conv(sp, st["sr0"], ws.fmap("hr_a64", B, H, W, 64), act=ACT_RELU)
        b = K.conv(a, st["sr1"], ws.fmap("hr_b64", B, H, W, 64), act=ACT_LEAKY, slope=0.01)
        a = K.conv(b, st["sr2"], ws.fmap("hr_a64", B, H, W, 64), act=ACT_LEAKY, slope=0.01)
        b = K.conv(a, st["sr3"], ws.fmap("hr_b64", B, H, W, 64), act=ACT_LEAKY, slope=0.01)
        a = K.conv(b, st["sr4"], ws.fmap("hr_a64", B, H, W, 64), act=ACT_LEAKY, slope=0.01)
        b = K.conv(a, st["cat0_sr"], ws.fmap("hr_b64", B, H, W, 64), bias=cb.t, bias_sn=25 * 64, bias_sc=64, cls_bw=2,
                   act=ACT_LEAKY, slope=0.01)
        a = K.conv(b, st["cat1"], ws.fmap("hr_a64", B, H, W, 64), act=ACT_LEAKY, slope=0.01)
        b = K.conv(a, st["cat2"], ws.fmap("hr_b64", B, H, W, 64))
        d49 = K.gap(b, ws.f32("v49", B, kc), kc)
        if self.debug is not None:
            self.debug["delta49_%d" % s] = d49.clone()
        out = ws.f32("kvec_b" if (s % 2 == 0) else "kvec_a", B, cond)
        return K.kernel_update(d49, kvec, out, self.ke, self.ko, True)

    def _sft(self, st, feats, kvec, B, h, w, s):
        ws, C, cond = self.ws, self.C, self.ko * self.ko
        cond_pad = K.round_up(cond, 64)
        fc = (s + 1) * C
        cc_pad = K.round_up(fc + cond, 64)
        small = K.broadcast_vec(kvec, ws.fmap("k3", B, 3, 3, cond_pad))
        res = {}
        for br in ("shift", "scale"):
            cb = K.conv(small, st["sft.%s.0k" % br], F32Map(ws.f32("k3bias%d" % cc_pad, B, 3, 3, cc_pad)))
            t = K.conv(feats, st["sft.%s.0f" % br], ws.fmap("sft_t%d" % cc_pad, B, h, w, cc_pad), bias=cb.t,
                       bias_sn=9 * cc_pad, bias_sc=cc_pad, cls_bw=1, act=ACT_LEAKY, slope=0.1)
            if br == "shift":
                res = K.conv(t, st["sft.shift.1"], ws.fmap("sft_shift%d" % fc, B, h, w, fc))
            else:
                return K.conv(t, st["sft.scale.1"], ws.fmap("sft_out%d" % fc, B, h, w, fc), act=ACT_SIGMOID, rm=feats, r1=res)
