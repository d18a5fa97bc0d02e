"""Shared generators / comparison helpers for the parity tests."""
import torch


def random_csr(M, N, avg_deg, seed, power_law=False, empty_rows=(), long_rows=(), device="cpu"):
    """Random CSR structure: sorted unique columns per row; optional power-law degrees, forced empty
    rows and forced long rows [(row, degree)]."""
    g = torch.Generator().manual_seed(seed)
    if power_law:
        u = torch.rand(M, generator=g).clamp_(min=1e-6)
        deg = (avg_deg / 3.0 * u.pow(-1 / 1.5)).floor().long().clamp_(max=N)
    else:
        deg = torch.poisson(torch.full((M,), float(avg_deg)), generator=g).long().clamp_(max=N)
    for r in empty_rows:
        deg[r] = 0
    for r, d in long_rows:
        deg[r] = min(d, N)
    rows, cols = [], []
    for m in range(M):
        d = int(deg[m])
        if d == 0:
            continue
        c = torch.randperm(N, generator=g)[:d].sort().values
        rows.append(torch.full((d,), m, dtype=torch.long))
        cols.append(c)
    row = torch.cat(rows) if rows else torch.empty(0, dtype=torch.long)
    col = torch.cat(cols) if cols else torch.empty(0, dtype=torch.long)
    rowptr = torch.zeros(M + 1, dtype=torch.long)
    rowptr[1:] = torch.cumsum(deg, 0)
    return row.to(device), rowptr.to(device), col.to(device)


def fast_random_csr(M, N, avg_deg, seed, device):
    """Large uniform random CSR generated on `device` (SURVEY §8d G2 recipe): E0 = avg_deg*M random
    (row, col) pairs, sorted + deduplicated."""
    g = torch.Generator(device=device).manual_seed(seed)
    E0 = int(avg_deg * M)
    row = torch.randint(M, (E0,), generator=g, device=device)
    col = torch.randint(N, (E0,), generator=g, device=device)
    key = torch.unique(row * N + col)  # sorted
    row, col = key // N, key % N
    rowptr = torch.zeros(M + 1, dtype=torch.long, device=device)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=M), 0)
    return row, rowptr, col


def rel_err_bound(out, ref, absmax):
    """max |out - ref| / max(|A||B| normaliser)"""
    return ((out.double() - ref.double()).abs() / absmax.double().clamp_min(1e-30)).max().item()
