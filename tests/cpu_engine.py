"""CPU stand-in for `cyclic_gps._engine` used ONLY by the tests of the multi-GPU host logic
(`cyclic_gps.distributed`) under gloo: the same `forward_sweep` / `backward_sweep` contract
(including the left-halo extension of the level kernels), built from the oracle's per-level
algebra.  It is a second, independent statement of the halo semantics documented in
include/crb200.h; the product never imports it."""
import torch

from oracle import cr_oracle as orc


class Pack:
    def __init__(self):
        self.levels = []          # per level: list over batch of dict(K,F,G,x,Gh)
        self.ms = []
        self.logdet = None
        self.mahal = None
        self.rest = None
        self.halo_out = None
        self.info = None

    def check(self):
        return None


def _sizes(n):
    ms = [n]
    while ms[-1] > 1:
        ms.append(ms[-1] // 2)
    return ms


def forward_sweep(R, O, y, *, keep_factors, want_logdet=True, nlevels=None, halo_O=None):
    B, n, l = R.shape[0], R.shape[1], R.shape[2]
    ms_all = _sizes(n)
    L = len(ms_all) if nlevels is None else min(nlevels, len(ms_all))
    p = Pack()
    p.ms = ms_all[:L]
    p.batch, p.ell, p.dtype, p.n = B, l, R.dtype, n
    ld = torch.zeros(B, dtype=torch.float64)
    mh = torch.zeros(B, dtype=torch.float64)
    rests, Rh_all, yh_all, Oh_all = [], [], [], []
    per_level = [[None] * B for _ in range(L)]
    for b in range(B):
        Rb, Ob, yb = R[b].clone(), O[b].clone() if n > 1 else R.new_zeros((0, l, l)), (y[b].clone() if y is not None else None)
        Oh = halo_O[b].clone() if halo_O is not None else None
        Rh = R.new_zeros((l, l))
        yh = R.new_zeros((l,))
        for k in range(L):
            m = Rb.shape[0]
            if m > 1:
                (mm, K, F, G), (Rn, On) = orc.level_step(Rb, Ob)
            else:
                K, F, G, Rn, On = orc._chol(Rb), R.new_zeros((0, l, l)), R.new_zeros((0, l, l)), R.new_zeros((0, l, l)), R.new_zeros((0, l, l))
            ld[b] += torch.log(torch.diagonal(K, dim1=-2, dim2=-1)).sum().double()
            xk = None
            yn = None
            if yb is not None:
                xk = torch.linalg.solve_triangular(K, yb[0::2].unsqueeze(-1), upper=False).squeeze(-1)
                mh[b] += (xk.double() ** 2).sum()
                yn = yb[1::2] - orc.bidiag_mv(F, G, xk) if m > 1 else R.new_zeros((0, l))
            Gh = None
            if Oh is not None:
                # virtual surviving node -1 coupled to row 0 by Oh = J_{0,-1}
                Gh = torch.linalg.solve_triangular(K[0], Oh, upper=False).mT          # O^T K^{-T}
                Rh = Rh - Gh @ Gh.mT
                if yb is not None:
                    yh = yh - Gh @ xk[0]
                Oh = -(F[0] @ Gh.mT) if m > 1 else R.new_zeros((l, l))
            per_level[k][b] = dict(K=K, F=F, G=G, x=xk, Gh=Gh)
            Rb, Ob, yb = Rn, On, yn
        rests.append((Rb, Ob, yb))
        Rh_all.append(Rh); yh_all.append(yh); Oh_all.append(Oh if Oh is not None else R.new_zeros((l, l)))
    p.levels = per_level
    p.logdet = 2.0 * ld if want_logdet else None
    p.mahal = mh if y is not None else None
    if L < len(ms_all):
        p.rest = (torch.stack([r[0] for r in rests]), torch.stack([r[1] for r in rests]),
                  torch.stack([r[2] for r in rests]) if y is not None else None)
    if halo_O is not None:
        p.halo_out = dict(Rh=torch.stack(Rh_all), yh=torch.stack(yh_all), O=torch.stack(Oh_all))
    return p


def backward_sweep(pack, *, sigma, w, xs=None, grad=None, top=None, halo=None, out=None):
    B, l = pack.batch, pack.ell
    L = len(pack.ms)
    Sd_all, So_all, w_all, Soh_all = [], [], [], []
    for b in range(B):
        if top is not None:
            Sd, So, wv = top[0][b], (top[1][b] if top[1] is not None else top[0].new_zeros((0, l, l))), top[2][b]
        else:
            Sd = So = wv = None
        Soh = halo["So"][b] if halo is not None else None
        for k in range(L - 1, -1, -1):
            lv = pack.levels[k][b]
            K, F, G, xk, Gh = lv["K"], lv["F"], lv["G"], lv["x"], lv["Gh"]
            E, o = K.shape[0], F.shape[0]
            Di = torch.linalg.inv(K)
            DtD = Di.mT @ Di
            if Sd is None:                    # deepest level of a complete sweep: m == 1
                Sd_new = DtD
                So_new = K.new_zeros((0, l, l))
                rhs = xk.clone()
            else:
                P = F @ Di[:o]
                Q = G @ Di[1:1 + G.shape[0]]
                mid, hi = orc.symtri_times_bidiag(-Sd, -So, P, Q)
                Se = DtD - orc.bidiag_t_bidiag_diag(P, Q, mid, hi)
                Sd_new, So_new = orc.interleave(Se, Sd), orc.interleave(mid, hi.mT)
                rhs = xk - orc.bidiag_tmv(F, G, wv)
            Soh_new = None
            if halo is not None:
                # contributions of the link between the virtual node -1 and even node 0
                Qh = Gh @ Di[0]
                Sdh, wh = halo["Sd"][b], halo["w"][b]
                cross = Soh.mT @ (F[0] @ Di[0]) if (Sd is not None and o > 0) else K.new_zeros((l, l))
                So_left = -(Sdh @ Qh + cross)                      # Sigma_{-1, 0}
                Sd_new = Sd_new.clone()
                extra = Qh.mT @ So_left
                if Sd is not None and o > 0:
                    # S_d[0] also gets -S~_o[-1] Q ; Sigma_even[0] gets -P^T (that extra)
                    P0 = F[0] @ Di[0]
                    add_mid = -(Soh @ Qh)
                    So_new = So_new.clone()
                    So_new[0] = So_new[0] + add_mid
                    Sd_new[0] = Sd_new[0] - P0.mT @ add_mid
                Sd_new[0] = Sd_new[0] - extra
                rhs = rhs.clone()
                rhs[0] = rhs[0] - Gh.mT @ wh
                Soh_new = So_left.mT                               # Sigma_{0,-1}
            w_even = torch.linalg.solve_triangular(K.mT, rhs.unsqueeze(-1), upper=True).squeeze(-1)
            wv = orc.interleave(w_even, wv) if Sd is not None else w_even
            Sd, So = Sd_new, So_new
            if halo is not None:
                Soh = Soh_new
        if grad is not None:
            gm, gd = float(grad[0][b]), float(grad[1][b])
            wl = halo["w"][b] if halo is not None else None
            gR = gd * Sd - gm * torch.einsum("ia,ib->iab", wv, wv)
            gO = 2 * gd * So - 2 * gm * torch.einsum("ia,ib->iab", wv[1:], wv[:-1])
            if halo is not None:
                Soh = 2 * gd * Soh - 2 * gm * torch.outer(wv[0], wl)
            Sd, So, wv = gR, gO, 2 * gm * wv
        Sd_all.append(Sd); So_all.append(So); w_all.append(wv)
        if halo is not None:
            Soh_all.append(Soh)
    res = (torch.stack(Sd_all), torch.stack(So_all), torch.stack(w_all))
    if out is not None:
        for dst, src in zip(out, res):
            if dst is not None and dst.numel():
                dst.copy_(src)
        res = tuple(out)
    if halo is not None:
        return res + (torch.stack(Soh_all),)
    return res
