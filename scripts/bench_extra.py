"""Secondary measurements of the hot path (not the driver bench): C3 SpMM_max fwd+bwd, C4 SpSpMM,
coalesce, csr2csc, SpMM backward, PCIe bandwidth. Prints one JSON object per line."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch
import pytorch_sparse_b200 as ts
from pytorch_sparse_b200 import ops
from util import fast_random_csr

dev = "cuda:0"
which = set(sys.argv[1:]) or {"pcie", "c2bw", "c3", "c4", "coalesce", "csc"}


def timeit(fn, steps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def out(**kw):
    print(json.dumps(kw), flush=True)


if "pcie" in which:
    h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms_up = timeit(lambda: d.copy_(h, non_blocking=True), 5, 1)
    ms_dn = timeit(lambda: h.copy_(d, non_blocking=True), 5, 1)
    out(what="pcie_pinned_256MB", h2d_gbs=256 / 1024 / ms_up * 1e3, d2h_gbs=256 / 1024 / ms_dn * 1e3)
    for ns in (2, 4, 8):
        streams = [torch.cuda.Stream() for _ in range(ns)]
        chunk = (256 << 20) // ns
        def multi():
            ev = torch.cuda.Event(); ev.record()
            for i, st in enumerate(streams):
                st.wait_event(ev)
                with torch.cuda.stream(st):
                    d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
            for st in streams:
                torch.cuda.current_stream().wait_stream(st)
        ms = timeit(multi, 5, 1)
        out(what=f"pcie_h2d_{ns}_streams", gbs=256 / 1024 / ms * 1e3)
    # simultaneous H2D + D2H
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h2 = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); d2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    def both():
        ev = torch.cuda.Event(); ev.record()
        s1.wait_event(ev); s2.wait_event(ev)
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    ms = timeit(both, 5, 1)
    out(what="pcie_bidirectional_256MB_each", ms=ms, agg_gbs=512 / 1024 / ms * 1e3)

if "overhead" in which:  # host-side launch cost of one small SpMM through the public API (no sync inside the loop)
    M = 10_000
    row, rowptr, col = fast_random_csr(M, M, 5, 1, dev)
    a = ts.SparseTensor(row=row, rowptr=rowptr, col=col, value=torch.rand(col.numel(), device=dev), sparse_sizes=(M, M),
                        is_sorted=True, trust_data=True)
    x = torch.randn(M, 32, device=dev)
    for _ in range(50):
        a @ x
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000):
        a @ x
    host_us = (time.perf_counter() - t0) / 2000 * 1e6
    torch.cuda.synchronize()
    t_dev = timeit(lambda: a @ x, 200, 10) * 1e3
    rowptr_, col_, val_ = a.csr()
    t0 = time.perf_counter()
    for _ in range(2000):
        ops.spmm_fw(rowptr_, col_, val_, x, "sum")
    torch.cuda.synchronize()
    raw_us = (time.perf_counter() - t0) / 2000 * 1e6
    out(what="c1_overhead", api_host_us_per_call=host_us, device_us_per_call=t_dev, ops_spmm_fw_us_per_call=raw_us)

if "e2e" in which:
    M = 1_000_000; F = 128
    row, rowptr, col = fast_random_csr(M, M, 16, 1, dev)
    val = (torch.rand(col.numel(), device=dev) + 0.5).bfloat16()
    x = torch.randn(M, F, device=dev).bfloat16()
    rp_h, col_h, val_h, x_h = [t.cpu().pin_memory() for t in (rowptr, col, val, x)]
    for _ in range(2):
        ops.spmm_fw_host(rp_h, col_h, val_h, x_h, "sum")
    t0 = time.perf_counter()
    for _ in range(5):
        o, _ = ops.spmm_fw_host(rp_h, col_h, val_h, x_h, "sum")
    ms = (time.perf_counter() - t0) * 1e3 / 5
    import os
    out(what="e2e_host_c2", chunks=os.environ.get("TSB200_HOST_CHUNKS", "default"), ms=ms)
    # split: pinned allocation vs the library call with a preallocated pinned output
    t0 = time.perf_counter()
    for _ in range(5):
        tmp = torch.empty(M, F, dtype=torch.bfloat16, pin_memory=True)
    alloc_ms = (time.perf_counter() - t0) * 1e3 / 5
    import ctypes
    from pytorch_sparse_b200._lib import lib
    o = torch.empty(M, F, dtype=torch.bfloat16, pin_memory=True)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    def raw():
        return lib.tsb200_spmm_fw_host(P(rp_h), P(col_h), P(val_h), P(x_h), P(o), None, 1, M, M, F, col_h.numel(), 3, 0)
    raw(); raw()
    t0 = time.perf_counter()
    for _ in range(5):
        rc = raw()
    raw_ms = (time.perf_counter() - t0) * 1e3 / 5
    out(what="e2e_split", pinned_alloc_256MB_ms=alloc_ms, lib_call_ms=raw_ms, rc=rc)

if "c2bw" in which:  # SpMM_sum fwd+bwd at C2 (value grad + dense grad)
    M = 1_000_000; F = 128
    row, rowptr, col = fast_random_csr(M, M, 16, 1, dev)
    E = col.numel()
    a = ts.SparseTensor(row=row, rowptr=rowptr, col=col, value=(torch.rand(E, device=dev) + 0.5).bfloat16(),
                        sparse_sizes=(M, M), is_sorted=True, trust_data=True).requires_grad_()
    x = torch.randn(M, F, device=dev).bfloat16().requires_grad_()
    go = torch.randn(M, F, device=dev).bfloat16()
    t_csc = timeit(lambda: (a.storage.clear_cache_(), a.storage.csr2csc()), 5, 1)
    a.storage.fill_cache_()
    t_f = timeit(lambda: a @ x.detach(), 20, 3)
    def fb():
        x.grad = None; a.storage.value().grad = None
        (a @ x).backward(go)
    t_fb = timeit(fb, 10, 2)
    t_vbw = timeit(lambda: ops.spmm_value_bw(row, rowptr, col, x.detach(), go, "sum"), 10, 2)
    out(what="c2_sum_bf16_F128", fwd_ms=t_f, fwd_bwd_ms=t_fb, value_bw_ms=t_vbw, csr2csc_colptr_ms=t_csc, E=E)

if "c3" in which:  # SpMM_max fwd + bwd, power-law, F=256 fp32 (BASELINE configs[2])
    import bench
    w = bench.WORKLOADS["c3"]
    rowptr, col, value, N = bench.gen_matrix(w, 0, 1)
    M, F = w["M"], w["F"]
    rowptr, col, value = rowptr.to(dev), col.to(dev), value.to(dev)
    E = col.numel()
    deg = rowptr[1:] - rowptr[:-1]
    a = ts.SparseTensor(rowptr=rowptr, col=col, value=value, sparse_sizes=(M, N), is_sorted=True, trust_data=True).requires_grad_()
    x = torch.randn(N, F, device=dev).requires_grad_()
    go = torch.randn(M, F, device=dev)
    t_f = timeit(lambda: a.matmul(x.detach(), "max"), 10, 2)
    def fb():
        x.grad = None; a.storage.value().grad = None
        a.matmul(x, "max").backward(go)
    t_fb = timeit(fb, 5, 2)
    out(what="c3_max_f32_F256_powerlaw", fwd_ms=t_f, fwd_bwd_ms=t_fb, E=E, max_deg=int(deg.max()), empty_rows=int((deg == 0).sum()),
        fwd_alg_gbs=bench.algorithmic_bytes(M, N, E, F, 4, True) / t_f / 1e6)

if "c4" in which:  # SpSpMM 256k x 256k, 32 nnz/row, fp32 (BASELINE configs[3])
    import os
    M = 262_144
    ra, rpa, ca = fast_random_csr(M, M, 32, 3, dev)
    rb, rpb, cb = fast_random_csr(M, M, 32, 4, dev)
    va = torch.randn(ca.numel(), device=dev); vb = torch.randn(cb.numel(), device=dev)
    from pytorch_sparse_b200._lib import lib
    from pytorch_sparse_b200.ops import _p, _stream, _workspace
    res = {}
    def run():
        res["c"] = ops.spspmm(rpa, ca, va, rpb, cb, vb, M, M, M, True)
    ref = None
    for mode in ("two_phase", "fused"):
        os.environ["TSB200_SPSPMM"] = mode
        torch.cuda.synchronize(); t0 = time.perf_counter(); run(); torch.cuda.synchronize(); t_first = (time.perf_counter() - t0) * 1e3
        t = timeit(run, 5, 2)   # two warm-ups: the caching allocator needs both output sets (old result alive while the new one is built)
        rp_c, r_c, c_c, v_c = res["c"]
        nnz = c_c.numel()
        if ref is None:
            ref = (rp_c.clone(), c_c.clone(), v_c.clone())
        else:  # both paths agree: structure bit-exact, values to rounding (accumulation order of duplicates)
            assert torch.equal(rp_c, ref[0]) and torch.equal(c_c, ref[1])
            assert torch.allclose(v_c, ref[2], rtol=1e-4, atol=1e-5)
        out(what="c4_spspmm_f32", mode=mode, ms=t, first_ms=t_first, nnz_a=ca.numel(), nnz_b=cb.numel(), nnz_c=nnz,
            out_gbs=nnz * 20 / t / 1e6, Gnnz_per_s=nnz / t / 1e6)
    # the kernels alone (no allocation, no nnz readback): C-ABI calls on preallocated outputs
    nws = lib.tsb200_spspmm_workspace_bytes(M, M, M, ca.numel(), cb.numel()); ws = _workspace(nws, torch.device(dev)); st = _stream(torch.device(dev))
    rp_tmp = torch.empty_like(rp_c)
    t_sym = timeit(lambda: lib.tsb200_spspmm_symbolic(_p(rpa), _p(ca), _p(rpb), _p(cb), M, M, M, ca.numel(), cb.numel(),
                                                     _p(rp_tmp), _p(ws), nws, None, st), 5, 1)
    assert torch.equal(rp_tmp, rp_c)
    t_num = timeit(lambda: lib.tsb200_spspmm_numeric(_p(rpa), _p(ca), _p(va), _p(rpb), _p(cb), _p(vb), M, M, M, ca.numel(),
                                                    cb.numel(), _p(rp_c), _p(r_c), _p(c_c), _p(v_c), 0, _p(ws), nws, st), 5, 1)
    products = int(((rpb[1:] - rpb[:-1])[ca]).sum())
    r2 = torch.empty(products, dtype=torch.long, device=dev); c2 = torch.empty_like(r2); v2 = torch.empty(products, device=dev)
    t_fused = timeit(lambda: lib.tsb200_spspmm_fused(_p(rpa), _p(ca), _p(va), _p(rpb), _p(cb), _p(vb), M, M, M, ca.numel(),
                                                    cb.numel(), _p(rp_tmp), _p(r2), _p(c2), _p(v2), products, 0, _p(ws), nws,
                                                    None, st), 5, 1)
    assert torch.equal(rp_tmp, rp_c) and torch.equal(c2[:nnz], c_c) and torch.equal(r2[:nnz], r_c)
    out(what="c4_spspmm_kernels", symbolic_ms=t_sym, numeric_ms=t_num, fused_ms=t_fused, numeric_out_gbs=nnz * 20 / t_num / 1e6,
        fused_out_gbs=nnz * 20 / t_fused / 1e6, products=products)
    del res, rp_c, r_c, c_c, v_c, r2, c2, v2, ref

if "coalesce" in which:
    M = N = 262_144
    g = torch.Generator(device=dev).manual_seed(5)
    E0 = 4_194_304
    row = torch.randint(M, (E0,), generator=g, device=dev); col = torch.randint(N, (E0,), generator=g, device=dev)
    perm = torch.randperm(2 * E0, generator=g, device=dev)
    row2, col2 = torch.cat([row, row])[perm], torch.cat([col, col])[perm]
    val = torch.randn(2 * E0, device=dev)
    t = timeit(lambda: ops.coalesce(row2, col2, val, M, N, "add"), 5, 1)
    t_sorted = timeit(lambda: ops.coalesce(*ops.coalesce(row2, col2, val, M, N, "add")[:2], None, M, N, "add"), 3, 1)
    def ref():
        key = row2 * N + col2
        k, inv = torch.unique(key, return_inverse=True)
        return torch.zeros(k.numel(), device=dev).index_add_(0, inv, val)
    t_torch = timeit(ref, 3, 1)
    out(what="coalesce_8.4M_shuffled_dup", ms=t, Mkeys_per_s=2 * E0 / t / 1e3, torch_unique_index_add_ms=t_torch)

if "csc" in which:
    M = 1_000_000
    row, rowptr, col = fast_random_csr(M, M, 16, 1, dev)
    t = timeit(lambda: ops.csr2csc(row, col, M, M, True, True), 5, 1)
    t_i2p = timeit(lambda: ops.ind2ptr(row, M), 10, 2)
    t_p2i = timeit(lambda: ops.ptr2ind(rowptr, col.numel()), 10, 2)
    t_torch = timeit(lambda: torch.sort(col * M + row)[1], 3, 1)
    out(what="format_16M", csr2csc_ms=t, ind2ptr_ms=t_i2p, ptr2ind_ms=t_p2i, torch_sort_linearised_ms=t_torch)
