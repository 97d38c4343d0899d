"""TEST DOUBLE -- a CPU (torch, fp64) stand-in for the seven CUDA stage calls of ``libsclip.so``.

It exists only so that the world_size > 1 choreography in ``synergy_clip_b200.ops`` (all-gather of the
normalised shards, merge of the column statistics, reduce-scatter of the column-role gradients, DDP
gradient scaling) can run under the gloo backend on a machine without a GPU.  It reads and writes the
same workspace sub-buffers, at the offsets ``sclip_plan`` reports, with the same meaning as the kernels
(``include/sclip.h``); tile-level partials are collapsed into tile 0.  It is never importable from the
package and is not a fallback: ``tests/test_multirank_cpu.py`` installs it explicitly.
"""
from __future__ import annotations

import math

import torch

KAPPA = 32768.0
PAIRS = ((0, 1), (1, 2), (2, 0))  # (row modality, column modality) of IT, TA, AI


class EmulatedBackend:
    allows_cpu = True

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _dims(ws):
        pb = ws.pb
        return pb.rows_local, pb.rows_global, pb.dim, pb.row_offset

    @staticmethod
    def _x3(ws):
        return ws.pb.math == 1

    def _xhat(self, ws):
        """Operands as the tensor cores see them: fp16 values (hi + lo in the split mode), fp64 here."""
        bl, bg, d, off = self._dims(ws)
        hi = ws.view(ws.lay.xhat, (3, bg, d), torch.float16).double()
        if self._x3(ws):
            return (hi + ws.view(ws.lay.xhat_lo, (3, bg, d), torch.float16).double()) / 256.0
        return hi

    # ------------------------------------------------------------------ forward
    def prologue(self, ws, img, txt, aud, t3=None, diag=False):
        bl, bg, d, off = self._dims(ws)
        hi = ws.view(ws.lay.xhat, (3, bg, d), torch.float16)
        inv = ws.view(ws.lay.inv_norm, (3, bl), torch.float32)
        for m, x in enumerate((img, txt, aud)):
            x = x.float()
            nrm = x.norm(dim=-1, keepdim=True)
            inv[m] = (1.0 / nrm).squeeze(-1)
            xh = x / nrm
            if self._x3(ws):
                scaled = xh * 256.0
                h = scaled.half()
                hi[m, off:off + bl] = h
                ws.view(ws.lay.xhat_lo, (3, bg, d), torch.float16)[m, off:off + bl] = (scaled - h.float()).half()
            else:
                hi[m, off:off + bl] = xh.half()
        ws.view(ws.lay.status, (4,), torch.int32).zero_()
        if diag:
            self.forward_diag(ws, t3)

    def _logits(self, ws, t3, p):
        bl, bg, d, off = self._dims(ws)
        xh = self._xhat(ws)
        r, c = PAIRS[p]
        cos = xh[r, off:off + bl] @ xh[c].T
        return cos, math.exp(float(t3[p])) * cos

    def forward_tiles(self, ws, t3):
        self.forward_tiles_cols(ws, t3, 7, 0, ws.lay.col_tiles)

    def forward_diag(self, ws, t3):
        bl, bg, d, off = self._dims(ws)
        xh = ws.view(ws.lay.xhat, (3, bg, d), torch.float16).double()
        dg = ws.view(ws.lay.diag_all, (3, bg), torch.float32)
        for p, (r, c) in enumerate(PAIRS):
            dg[p, off:off + bl] = (math.exp(float(t3[p])) * (xh[r, off:off + bl] * xh[c, off:off + bl]).sum(-1)).float()

    def forward_tiles_cols(self, ws, t3, pair_mask, tile_lo, tile_hi, stash=False, wrap=False, max_sms=0,
                           wait_epoch=None):
        """Column tiles [tile_lo, tile_hi) of the pairs in pair_mask; row partials land in slot `tile_lo`."""
        bl, bg, d, off = self._dims(ws)
        lay = ws.lay
        if tile_hi <= tile_lo:
            return
        row_part = ws.view(lay.row_part, (3, 2 * lay.col_tiles, bl), torch.float32)
        col_part = ws.view(lay.col_part, (3, lay.row_tiles, bg), torch.float32)
        tile_ref = ws.view(lay.tile_ref, (3, lay.row_tiles, lay.col_tiles), torch.float32)
        diag = ws.view(lay.diag, (3, bl), torch.float32)
        c0, c1 = tile_lo * 256, min(tile_hi * 256, bg)
        for p in range(3):
            if not pair_mask & (1 << p):
                continue
            _, logits = self._logits(ws, t3, p)
            # one reference for the whole pair (the kernels use one per tile; any consistent choice merges the same)
            ref = 0.0 if math.exp(float(t3[p])) < 64.0 else float(math.exp(float(t3[p])))
            tile_ref[p] = ref
            e = torch.exp(logits[:, c0:c1] - ref)
            row_part[p, 2 * tile_lo + 1:2 * tile_hi] = 0.0
            row_part[p, 2 * tile_lo] = e.sum(1).float()
            col_part[p, :, c0:c1] = 0.0
            col_part[p, 0, c0:c1] = e.sum(0).float()
            if c0 <= off < c1:
                diag[p] = logits[torch.arange(bl), off + torch.arange(bl)].float()
            if stash:
                dg = ws.view(lay.diag_all, (3, bg), torch.float32).double()[p]
                st = torch.exp(logits[:, c0:c1] - 0.5 * (dg[off:off + bl, None] + dg[None, c0:c1])) / 16.0
                ws.view(lay.grad_tiles, (3, bl, lay.ld_g), torch.float16)[p, :, c0:c1] = st.clamp(max=65504.0).half()

    def forward_reduce(self, ws):
        bl, bg, d, off = self._dims(ws)
        lay = ws.lay
        row_part = ws.view(lay.row_part, (3, 2 * lay.col_tiles, bl), torch.float32).double()
        col_part = ws.view(lay.col_part, (3, lay.row_tiles, bg), torch.float32).double()
        ref = ws.view(lay.tile_ref, (3, lay.row_tiles, lay.col_tiles), torch.float32).double()[:, 0, 0]
        rsum, csum = row_part.sum(1), col_part.sum(1)
        ws.view(lay.lse_row, (3, bl), torch.float32).copy_(ref[:, None] + rsum.log())
        ws.view(lay.lse_col_local, (3, bg), torch.float32).copy_(ref[:, None] + csum.log())
        ws.view(lay.row_inv, (3, bl), torch.float32).copy_(1.0 / rsum)
        ws.view(lay.col_sum_local, (3, bg), torch.float32).copy_(csum)

    def forward_loss(self, ws, col_lse_all, loss3):
        bl, bg, d, off = self._dims(ws)
        lay = ws.lay
        lse_col = ws.view(lay.lse_col, (3, bg), torch.float32)
        col_inv = ws.view(lay.col_inv, (3, bg), torch.float32)
        if col_lse_all is None:
            lse_col.copy_(ws.view(lay.lse_col_local, (3, bg), torch.float32))
            col_inv.copy_(1.0 / ws.view(lay.col_sum_local, (3, bg), torch.float32))
        else:
            merged = torch.logsumexp(col_lse_all.double(), dim=0)
            lse_col.copy_(merged)
            col_inv.copy_(torch.exp(-merged))
        lse_row = ws.view(lay.lse_row, (3, bl), torch.float32).double()
        diag = ws.view(lay.diag, (3, bl), torch.float32).double()
        part = ((lse_row - diag).sum(1) + (lse_col.double()[:, off:off + bl] - diag).sum(1)) / (2.0 * bg)
        ws.view(lay.loss_part, (3,), torch.float32).copy_(part)
        loss3.copy_(part)

    # ------------------------------------------------------------------ backward
    @staticmethod
    def _coeffs(t3, g3):
        sg = [math.exp(float(t)) * float(g) for t, g in zip(t3, g3)]
        mx = max(abs(v) for v in sg)
        return mx, [v / mx if mx > 0 else 0.0 for v in sg]

    def backward_tiles(self, ws, t3, g3):
        bl, bg, d, off = self._dims(ws)
        lay = ws.lay
        _, cps = self._coeffs(t3, g3)
        g_hi = ws.view(lay.grad_tiles, (3, bl, lay.ld_g), torch.float16)
        dt_part = ws.view(lay.dt_part, (3, lay.row_tiles * lay.col_tiles), torch.float32).zero_()
        lse_row = ws.view(lay.lse_row, (3, bl), torch.float32).double()
        lse_col = ws.view(lay.lse_col, (3, bg), torch.float32).double()
        for p in range(3):
            cos, logits = self._logits(ws, t3, p)
            soft = torch.exp(logits - lse_row[p][:, None]) + torch.exp(logits - lse_col[p][None, :])
            gp = 0.5 * KAPPA * cps[p] * soft
            gp[torch.arange(bl), off + torch.arange(bl)] -= KAPPA * cps[p]
            dt_part[p, 0] = float((gp * cos).sum())
            h = gp.half()
            g_hi[p, :, :bg] = h
            if self._x3(ws):
                ws.view(lay.grad_tiles_lo, (3, bl, lay.ld_g), torch.float16)[p, :, :bg] = (gp - h.double()).half()

    def backward_scale(self, ws, t3, g3):
        bl, bg, d, off = self._dims(ws)
        lay = ws.lay
        _, cps = self._coeffs(t3, g3)
        g = ws.view(lay.grad_tiles, (3, bl, lay.ld_g), torch.float16)
        dg = ws.view(lay.diag_all, (3, bg), torch.float32).double()
        lse_row = ws.view(lay.lse_row, (3, bl), torch.float32).double()
        lse_col = ws.view(lay.lse_col, (3, bg), torch.float32).double()
        for p in range(3):
            k8 = 8.0 * KAPPA * cps[p]
            hr, hc = 0.5 * dg[p, off:off + bl], 0.5 * dg[p]
            fac = (k8 * torch.exp(hr - lse_row[p]))[:, None] * torch.exp(hc)[None, :] + \
                torch.exp(hr)[:, None] * (k8 * torch.exp(hc - lse_col[p]))[None, :]
            gp = g[p, :, :bg].double() * fac
            gp[torch.arange(bl), off + torch.arange(bl)] -= KAPPA * cps[p]  # the identity term, before the fp16 rounding
            g[p, :, :bg] = gp.half()

    def _gprime(self, ws):
        bl, bg, d, off = self._dims(ws)
        lay = ws.lay
        g = ws.view(lay.grad_tiles, (3, bl, lay.ld_g), torch.float16)[:, :, :bg].double()
        if self._x3(ws):
            g = g + ws.view(lay.grad_tiles_lo, (3, bl, lay.ld_g), torch.float16)[:, :, :bg].double()
        return g

    def backward_gemms(self, ws, t3, g3):
        self.backward_gemms_role(ws, t3, g3, 0)

    def backward_gemms_role(self, ws, t3, g3, role, max_sms=0, convert=False):
        bl, bg, d, off = self._dims(ws)
        lay = ws.lay
        mx, _ = self._coeffs(t3, g3)
        alpha = mx / (KAPPA * bg)
        xh = self._xhat(ws)
        g = self._gprime(ws)
        row_out = ws.view(lay.dxhat_row, (3, bl, d), torch.float32)
        for m in range(3):
            pr, pc = m, (m + 2) % 3
            acc = g[pr] @ xh[PAIRS[pr][1]]                              # row role, k = global row
            col_role = g[pc].T @ xh[PAIRS[pc][0], off:off + bl]          # column role, k = local row -> (bg, d)
            if ws.pb.world == 1:
                row_out[m] = (alpha * (acc + col_role)).float()
            else:
                if role in (0, 2):
                    row_out[m] = (alpha * acc).float()
                if role in (0, 1):
                    ws.view(lay.dxhat_col, (3, bg, d), torch.float32)[m] = (alpha * col_role).float()

    def backward_finish(self, ws, img, txt, aud, t3, g3, col, mult, dimg, dtxt, daud, out_f32, dt3, stashed=False):
        bl, bg, d, off = self._dims(ws)
        lay = ws.lay
        inv = ws.view(lay.inv_norm, (3, bl), torch.float32).double()
        row_out = ws.view(lay.dxhat_row, (3, bl, d), torch.float32).double()
        xh16 = ws.view(lay.xhat, (3, bg, d), torch.float16).double()[:, off:off + bl]
        dots = []
        for m, (x, out) in enumerate(zip((img, txt, aud), (dimg, dtxt, daud))):
            dd = row_out[m] + (col[m].double() if col is not None else 0.0)
            xh = x.double() * inv[m][:, None]
            dots.append(float((xh * dd).sum()))
            dx = (dd - xh * (xh * dd).sum(-1, keepdim=True)) * inv[m][:, None] * mult
            out.copy_(dx.to(out.dtype))
        if stashed:
            dt = [0.5 * (dots[p] + dots[(p + 1) % 3] - dots[(p + 2) % 3]) * mult for p in range(3)]
            dt3.copy_(torch.tensor(dt, dtype=torch.float32))
        else:
            mx, _ = self._coeffs(t3, g3)
            dt_part = ws.view(lay.dt_part, (3, lay.row_tiles * lay.col_tiles), torch.float32).double()
            dt3.copy_((dt_part.sum(1) * mx / (KAPPA * bg) * mult).float())
