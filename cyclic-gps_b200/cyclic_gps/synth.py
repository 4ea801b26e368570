"""Synthetic LEG posterior-precision blocks on the GPU, for benchmarks and full-size tests
(SURVEY 8(d)).  Restates ``LEGFamily.compute_PEG_precision`` / ``compute_posterior_precision``
of the reference (cyclic_gps/models.py:181-239, 254-268) in batched torch ops; always built in
fp64 from fp64 time gaps and cast at the end (fp32 time stamps cannot resolve unit gaps at
n >= 1e7).  Input generation only -- not part of the timed hot path."""
import torch


def leg_params(rank: int, seed: int = 0, device="cpu"):
    """(G, B, LL^T) exactly as a fresh ``LEGFamily(rank, obs_dim=1)`` builds them
    (models.py:93-121, 152-166): N = I, R = strictly-lower part of 0.2 (A - A^T), B = 0.5/sqrt(rank),
    Lambda = softplus(0.1)."""
    gen = torch.Generator().manual_seed(seed)
    eye = torch.eye(rank, dtype=torch.float64)
    A = torch.randn((rank, rank), generator=gen, dtype=torch.float64)
    Rm = torch.tril((A - A.T) * 0.2, diagonal=-1)
    B = torch.full((1, rank), 0.5 / rank ** 0.5, dtype=torch.float64)
    lam = torch.nn.functional.softplus(torch.tensor([[0.1]], dtype=torch.float64))
    G = eye + Rm - Rm.T + 1e-5 * eye
    LLT = lam @ lam.T + 1e-9 * torch.eye(1, dtype=torch.float64)
    return G.to(device), B.to(device), LLT.to(device)


def leg_precision_blocks(gaps, G, B, LLT, dtype, chunk: int = 1 << 20):
    """gaps (..., n-1) fp64 -> Rs (..., n, l, l), Os (..., n-1, l, l) of
    K = Sigma^{-1} + B^T (LL^T)^{-1} B, in `dtype`.  Processed in chunks of gaps to bound memory."""
    lead, nm1 = gaps.shape[:-1], gaps.shape[-1]
    l = G.shape[0]
    flat = gaps.reshape(-1, nm1)
    S = flat.shape[0]
    dev = gaps.device
    eye = torch.eye(l, dtype=torch.float64, device=dev)
    shift = eye + B.T @ torch.linalg.solve(LLT, B)
    Rs = torch.empty((S, nm1 + 1, l, l), dtype=dtype, device=dev)
    Os = torch.empty((S, nm1, l, l), dtype=dtype, device=dev)
    evals, evecs = torch.linalg.eig(G.cpu())                   # G is tiny; eigendecompose once (model_utils.py:12-29)
    evals, evecs = evals.to(dev), evecs.to(dev)
    evecs_inv = torch.linalg.inv(evecs)
    for s in range(S):
        Rs[s] = shift.to(dtype)
        for a in range(0, nm1, chunk):
            d = flat[s, a:a + chunk]
            A = ((evecs.unsqueeze(0) * torch.exp(-0.5 * d.reshape(-1, 1, 1) * evals.reshape(1, 1, -1))) @ evecs_inv).real
            At = A.transpose(1, 2)
            right = torch.linalg.solve(eye - A @ At, A)          # (I - A A^T)^{-1} A
            left = torch.linalg.solve(eye - At @ A, At)          # (I - A^T A)^{-1} A^T
            Os[s, a:a + chunk] = (-right).to(dtype)
            Rs[s, a + 1:a + chunk + 1] += (A @ left).to(dtype)   # contribution of gap i to row i+1
            Rs[s, a:a + d.shape[0]] += (At @ right).to(dtype)    # contribution of gap i to row i
    return Rs.reshape(*lead, nm1 + 1, l, l), Os.reshape(*lead, nm1, l, l)


def gaps_for_rows(lo: int, hi: int, n: int, seed: int, device, block: int = 1 << 20):
    """Deterministic gap d_j (between rows j and j+1) for j in [lo-1, hi-1], identical on every rank:
    gaps are drawn per block of `block` indices from a generator seeded by (seed, block index).
    Non-existent gaps (j = -1 and j = n-1) are returned as +inf (rows 0 and n-1 have one neighbour)."""
    j0, j1 = lo - 1, hi - 1                      # inclusive range of gap indices
    out = torch.full((j1 - j0 + 1,), float("inf"), dtype=torch.float64, device=device)
    a, b = max(j0, 0), min(j1, n - 2)
    if b >= a:
        for blk in range(a // block, b // block + 1):
            gen = torch.Generator(device=device).manual_seed(seed * 1000003 + blk)
            vals = -torch.log(torch.rand(block, generator=gen, dtype=torch.float64, device=device)) + 0.01
            s, e = max(a, blk * block), min(b, (blk + 1) * block - 1)
            out[s - j0:e - j0 + 1] = vals[s - blk * block:e - blk * block + 1]
    return out


def leg_precision_rows(gaps_ext, G, B, LLT, dtype, chunk: int = 1 << 20):
    """Rows [lo, hi) of the LEG posterior precision in the `Oprev` convention of
    cyclic_gps.distributed: gaps_ext[j] is the gap between rows lo-1+j and lo+j (length n_loc+1, +inf
    where the neighbour does not exist).  Returns R (n_loc,l,l), Oprev (n_loc,l,l) with
    Oprev[i] = J_{lo+i, lo+i-1} (zero for a non-existent predecessor)."""
    n_loc = gaps_ext.shape[0] - 1
    l = G.shape[0]
    dev = gaps_ext.device
    eye = torch.eye(l, dtype=torch.float64, device=dev)
    shift = eye + B.T @ torch.linalg.solve(LLT, B)
    R = torch.empty((n_loc, l, l), dtype=dtype, device=dev)
    Oprev = torch.empty((n_loc, l, l), dtype=dtype, device=dev)
    evals, evecs = torch.linalg.eig(G.cpu())
    evals, evecs = evals.to(dev), evecs.to(dev)
    evecs_inv = torch.linalg.inv(evecs)
    for a in range(0, n_loc, chunk):
        b = min(a + chunk, n_loc)
        d = gaps_ext[a:b + 1]                                     # gaps a-1 .. b-1 relative to local rows
        finite = torch.isfinite(d)
        dd = torch.where(finite, d, torch.ones_like(d))
        A = ((evecs.unsqueeze(0) * torch.exp(-0.5 * dd.reshape(-1, 1, 1) * evals.reshape(1, 1, -1))) @ evecs_inv).real
        At = A.transpose(1, 2)
        fwd = torch.linalg.solve(eye - A @ At, A)
        bwd = torch.linalg.solve(eye - At @ A, At)
        mask = finite.reshape(-1, 1, 1).to(torch.float64)
        from_prev = (A @ bwd) * mask                              # gap j adds this to row j+1
        to_next = (At @ fwd) * mask                               # gap j adds this to row j
        R[a:b] = (shift + from_prev[:-1] + to_next[1:]).to(dtype)
        Oprev[a:b] = (-(fwd * mask)[:-1]).to(dtype)
    return R, Oprev


def x_for_rows(lo: int, hi: int, ell: int, seed: int, device, dtype, block: int = 1 << 20):
    """Right-hand side rows [lo, hi) of ONE global series, identical whatever the partition: rows are drawn per block
    of `block` global indices from a generator seeded by (seed, block index) -- the same scheme as `gaps_for_rows`."""
    out = torch.empty((max(hi - lo, 0), ell), dtype=dtype, device=device)
    if hi <= lo:
        return out
    for blk in range(lo // block, (hi - 1) // block + 1):
        gen = torch.Generator(device=device).manual_seed(seed * 7919 + blk)
        vals = torch.randn((block, ell), generator=gen, dtype=torch.float32, device=device)
        s, e = max(lo, blk * block), min(hi, (blk + 1) * block)
        out[s - lo:e - lo] = vals[s - blk * block:e - blk * block].to(dtype)
    return out
