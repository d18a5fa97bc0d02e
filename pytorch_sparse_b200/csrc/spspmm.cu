// spspmm.cu — CSR x CSR -> CSR/COO sparse-sparse matmul for sm_100a.
//
// Replaces spspmm_sum -> torch.sparse.mm (torch_sparse/matmul.py:94-111; CPU: ATen sparse_matmul,
// CUDA: cuSPARSE SpGEMM), keeping its observable contract: output sorted by (row, col), unique,
// structural (numerical zeros are kept).
//
// Two-phase row-wise Gustavson with a per-CTA shared-memory BITMAP accumulator:
//   * a CTA owns one output row at a time (rows are pulled from a global atomic counter);
//   * every product a_ik * b_kj sets bit j of a bitmap over a window of <= 2^19 columns
//     (64 KB); a 2-level summary bitmap remembers which 32-bit words were touched, so counting,
//     ranking and clearing only visit touched words;
//   * popcounts of the bitmap give (symbolic) the row's nnz and (numeric) the rank of every
//     column => the output row is produced ALREADY SORTED by column, with no per-row sort and
//     no hash probing; values are accumulated with shared-memory atomics at their rank
//     (global atomics for rows wider than the shared accumulator);
//   * matrices wider than one window are processed window by window (B rows are column-sorted,
//     so only windows between the row's smallest and largest product column are visited).
// Everything is HBM-bound on the 16+s bytes per output nonzero that must be written.
#include <cstdlib>

#include <cub/cub.cuh>

#include "common.cuh"

namespace tsb {

constexpr int kSpThreads = 256;
constexpr int kAccCap = 4096;         // shared accumulator entries per window
constexpr int kMaxWindowBits = 1 << 19;

struct SpParams {
  const int64_t* rowptr_a; const int64_t* col_a; const void* val_a;
  const int64_t* rowptr_b; const int64_t* col_b; const void* val_b;
  int64_t M, Kd, N;
  int64_t* counts;          // symbolic: per-row nnz (written at rowptr_c + 1)
  const int64_t* rowptr_c;  // numeric
  int64_t* row_c; int64_t* col_c; void* val_c;
  unsigned int* counter;
  int window_bits;  // power of two, >= 1024
  int acc_cap;      // entries of the shared-memory value accumulator (0: accumulate with global atomics)
};

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, x, off);
    if (lane >= off) x += t;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (kSpThreads / 32) ? s_warp[lane] : 0;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, off);
      if (lane >= off) w += t;
    }
    if (lane < (kSpThreads / 32)) s_warp[lane] = w;  // inclusive warp totals
  }
  __syncthreads();
  const int warp_off = warp ? s_warp[warp - 1] : 0;
  total = s_warp[kSpThreads / 32 - 1];
  const int r = warp_off + x - v;
  __syncthreads();
  return r;
}

// Per output row the CTA first stages the A-row metadata in shared memory (for each a_ik: start and
// length of row k of B, and a_ik), ONCE; the product walks (mark, accumulate) then run warp-per-A-entry
// with lanes striding the B row: the only global loads of a walk are the independent, coalesced reads
// of B's column (and value) arrays — one L2/DRAM latency per walk instead of one per A entry.
constexpr int kABatch = kSpThreads;  // A entries staged per batch

template <bool NUMERIC, typename T>
__global__ void __launch_bounds__(kSpThreads) spspmm_kernel(const SpParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int WW = p.window_bits >> 5;  // bitmap words
  const int NSW = WW >> 5;            // summary words
  uint32_t* bitmap = (uint32_t*)smem;
  uint32_t* summary = bitmap + WW;
  uint32_t* base = summary + NSW;                  // NUMERIC only
  uint16_t* pre16 = (uint16_t*)(base + NSW);       // NUMERIC only
  T* acc = (T*)(pre16 + WW);                       // NUMERIC only (16 B aligned: WW*2 is a multiple of 64)
  __shared__ int s_warp[kSpThreads / 32];
  __shared__ unsigned int s_row;
  __shared__ long long s_min, s_max;
  __shared__ int64_t s_bs[kABatch];
  __shared__ int s_len[kABatch];
  __shared__ T s_av[NUMERIC ? kABatch : 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARP = kSpThreads / 32;
  for (int i = tid; i < WW + NSW; i += kSpThreads) bitmap[i] = 0;  // summary is contiguous after bitmap
  __syncthreads();

  const T* va = (const T*)p.val_a;
  const T* vb = (const T*)p.val_b;
  const int64_t W = p.window_bits;
  const bool multi_window = p.N > W;

  while (true) {
    if (tid == 0) s_row = atomicAdd(p.counter, 1u);
    __syncthreads();
    const int64_t i = s_row;
    const int64_t a_s = p.rowptr_a[i < p.M ? i : 0], a_e = p.rowptr_a[i < p.M ? i + 1 : 0];
    __syncthreads();
    if (i >= p.M) break;
    int64_t win_lo = 0, win_hi = 0;  // window index range [win_lo, win_hi]
    if (multi_window) {
      // B rows are column-sorted (SparseStorage invariant): first/last entry bound the row's columns
      if (tid == 0) { s_min = 0x7fffffffffffffffLL; s_max = -1; }
      __syncthreads();
      long long mn = 0x7fffffffffffffffLL, mx = -1;
      for (int64_t a = a_s + tid; a < a_e; a += kSpThreads) {
        const int64_t k = p.col_a[a];
        const int64_t bs = p.rowptr_b[k], be = p.rowptr_b[k + 1];
        if (be > bs) {
          mn = min(mn, (long long)p.col_b[bs]);
          mx = max(mx, (long long)p.col_b[be - 1]);
        }
      }
      if (mx >= 0) { atomicMin(&s_min, mn); atomicMax(&s_max, mx); }
      __syncthreads();
      if (s_max < 0) { win_lo = 1; win_hi = 0; }
      else { win_lo = s_min / W; win_hi = s_max / W; }
      __syncthreads();
    }
    const bool single_batch = (a_e - a_s) <= kABatch;
    int64_t done = 0;  // nnz of this row emitted by previous windows
    const int64_t out0 = NUMERIC ? p.rowptr_c[i] : 0;
    bool staged = false;

    for (int64_t win = win_lo; win <= win_hi; win++) {
      const int64_t wlo = win * W, whi = wlo + W;
      // ---- mark ----
      for (int64_t ab = a_s; ab < a_e; ab += kABatch) {
        const int na = (int)min((int64_t)kABatch, a_e - ab);
        if (!(single_batch && staged)) {
          if (tid < na) {
            const int64_t k = p.col_a[ab + tid];
            const int64_t bs = p.rowptr_b[k];
            s_bs[tid] = bs;
            s_len[tid] = (int)(p.rowptr_b[k + 1] - bs);
            if (NUMERIC) s_av[tid] = va ? va[ab + tid] : (T)1;
          }
          __syncthreads();
          staged = true;
        }
        for (int e = warp; e < na; e += NWARP) {
          const int64_t bs = s_bs[e];
          const int len = s_len[e];
          for (int f = lane; f < len; f += 32) {
            const int64_t c = p.col_b[bs + f];
            if (c >= wlo && c < whi) {
              const uint32_t cc = (uint32_t)(c - wlo);
              const uint32_t old = atomicOr(&bitmap[cc >> 5], 1u << (cc & 31));
              if (old == 0) atomicOr(&summary[cc >> 10], 1u << ((cc >> 5) & 31));
            }
          }
        }
        __syncthreads();
      }

      if (!NUMERIC) {
        int cnt = 0;
        for (int sw = tid; sw < NSW; sw += kSpThreads) {
          uint32_t m = summary[sw];
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int w = (sw << 5) + b;
            cnt += __popc(bitmap[w]);
            bitmap[w] = 0;
          }
          summary[sw] = 0;
        }
        int total;
        block_exclusive_scan(cnt, s_warp, total);
        done += total;
      } else {
        // ---- rank: prefix of popcounts over touched words ----
        const int spt = (NSW + kSpThreads - 1) / kSpThreads;
        const int sw0 = tid * spt, sw1 = min(NSW, sw0 + spt);
        int mine = 0;
        for (int sw = sw0; sw < sw1; sw++) {
          uint32_t m = summary[sw];
          int run = 0;
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int w = (sw << 5) + b;
            pre16[w] = (uint16_t)run;
            run += __popc(bitmap[w]);
          }
          base[sw] = run;
          mine += run;
        }
        int wc;
        int off = block_exclusive_scan(mine, s_warp, wc);
        for (int sw = sw0; sw < sw1; sw++) {
          const int t = base[sw];
          base[sw] = off;
          off += t;
        }
        const bool use_smem_acc = p.val_c != nullptr && wc <= p.acc_cap;
        const int64_t obase = out0 + done;
        if (p.val_c) {
          if (use_smem_acc) for (int q = tid; q < wc; q += kSpThreads) acc[q] = (T)0;
          else for (int q = tid; q < wc; q += kSpThreads) ((T*)p.val_c)[obase + q] = (T)0;
        }
        __syncthreads();
        // ---- emit columns (already sorted) ----
        for (int sw = tid; sw < NSW; sw += kSpThreads) {
          uint32_t m = summary[sw];
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int w = (sw << 5) + b;
            uint32_t bits = bitmap[w];
            int64_t pos = obase + base[sw] + pre16[w];
            while (bits) {
              const int bb = __ffs(bits) - 1;
              bits &= bits - 1;
              p.col_c[pos] = wlo + ((int64_t)w << 5) + bb;
              if (p.row_c) p.row_c[pos] = i;
              pos++;
            }
          }
        }
        // ---- accumulate values at their rank ----
        if (p.val_c) {
          for (int64_t ab = a_s; ab < a_e; ab += kABatch) {
            const int na = (int)min((int64_t)kABatch, a_e - ab);
            if (!single_batch) {
              __syncthreads();
              if (tid < na) {
                const int64_t k = p.col_a[ab + tid];
                const int64_t bs = p.rowptr_b[k];
                s_bs[tid] = bs;
                s_len[tid] = (int)(p.rowptr_b[k + 1] - bs);
                s_av[tid] = va ? va[ab + tid] : (T)1;
              }
              __syncthreads();
            }
            for (int e = warp; e < na; e += NWARP) {
              const int64_t bs = s_bs[e];
              const int len = s_len[e];
              const T av = s_av[e];
              for (int f = lane; f < len; f += 32) {
                const int64_t c = p.col_b[bs + f];
                if (c >= wlo && c < whi) {
                  const uint32_t cc = (uint32_t)(c - wlo);
                  const uint32_t w = cc >> 5;
                  const int rank = base[w >> 5] + pre16[w] + __popc(bitmap[w] & ((1u << (cc & 31)) - 1u));
                  const T pv = av * (vb ? vb[bs + f] : (T)1);
                  if (use_smem_acc) atomicAdd(&acc[rank], pv);
                  else atomicAdd(((T*)p.val_c) + obase + rank, pv);
                }
              }
            }
          }
          __syncthreads();
          if (use_smem_acc) for (int q = tid; q < wc; q += kSpThreads) ((T*)p.val_c)[obase + q] = acc[q];
        } else {
          __syncthreads();
        }
        // ---- clear touched words ----
        for (int sw = tid; sw < NSW; sw += kSpThreads) {
          uint32_t m = summary[sw];
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            bitmap[(sw << 5) + b] = 0;
          }
          summary[sw] = 0;
        }
        done += wc;
      }
      __syncthreads();
    }
    if (!NUMERIC && tid == 0) p.counts[i] = done;
  }
}

__global__ void sp_finish_kernel(int64_t* rowptr_c, int64_t M, int64_t* nnz_dev) {
  rowptr_c[0] = 0;
  (void)M;
  (void)nnz_dev;
}
__global__ void sp_copy_last_kernel(const int64_t* rowptr_c, int64_t M, int64_t* nnz_dev) { *nnz_dev = rowptr_c[M]; }

static int window_bits_for(int64_t N) {
  int64_t w = 1024;
  while (w < N && w < kMaxWindowBits) w <<= 1;
  return (int)w;
}
static size_t sp_smem_bytes(int window_bits, bool numeric, size_t elem, int acc_cap) {
  const size_t WW = (size_t)window_bits >> 5, NSW = WW >> 5;
  size_t b = WW * 4 + NSW * 4;
  if (numeric) b += NSW * 4 + WW * 2 + (size_t)acc_cap * elem;
  return b;
}

struct SpLayout { size_t scalars, cub, total, cub_bytes; };
static SpLayout sp_layout(int64_t M) {
  SpLayout L;
  size_t off = 0;
  L.scalars = off; off += 256;
  size_t tb = 0;
  cub::DeviceScan::InclusiveSum(nullptr, tb, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(M > 0 ? M : 1),
                                (cudaStream_t)0);
  L.cub_bytes = tb;
  L.cub = off; off += align_up(tb, 256);
  L.total = off;
  return L;
}

template <bool NUMERIC, typename T> static int sp_launch(const SpParams& p, cudaStream_t st) {
  const size_t smem = sp_smem_bytes(p.window_bits, NUMERIC, sizeof(T), p.acc_cap);
  auto* k = spspmm_kernel<NUMERIC, T>;
  TSB_CUDA_TRY(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  TSB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, kSpThreads, smem));
  if (per_sm < 1) per_sm = 1;
  int dev = 0, sms = kNumSMs;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t grid = (int64_t)per_sm * sms;
  if (grid > p.M) grid = p.M;
  if (grid < 1) grid = 1;
  k<<<(int)grid, kSpThreads, smem, st>>>(p);
  TSB_LAUNCH_CHECK();
  return 0;
}

}  // namespace tsb

using namespace tsb;

extern "C" size_t tsb200_spspmm_workspace_bytes(int64_t M, int64_t Kd, int64_t N, int64_t nnz_a, int64_t nnz_b) {
  (void)Kd; (void)N; (void)nnz_a; (void)nnz_b;
  if (M < 0) return 0;
  return sp_layout(M).total;
}

extern "C" int tsb200_spspmm_symbolic(const int64_t* rowptr_a, const int64_t* col_a, const int64_t* rowptr_b,
                                      const int64_t* col_b, int64_t M, int64_t Kd, int64_t N, int64_t nnz_a,
                                      int64_t nnz_b, int64_t* rowptr_c, void* workspace, size_t workspace_bytes,
                                      int64_t* nnz_c_host, void* stream) {
  if (M < 0 || Kd < 0 || N < 0 || nnz_a < 0 || nnz_b < 0 || !rowptr_c) return TSB200_ERR_INVALID_ARG;
  if (M >= ((int64_t)1 << 31)) return TSB200_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const SpLayout L = sp_layout(M);
  if (!workspace || workspace_bytes < L.total) return TSB200_ERR_WORKSPACE;
  char* ws = (char*)workspace;
  TSB_CUDA_TRY(cudaMemsetAsync(ws + L.scalars, 0, 256, st));
  TSB_CUDA_TRY(cudaMemsetAsync(rowptr_c, 0, (size_t)(M + 1) * 8, st));
  if (M == 0 || nnz_a == 0 || nnz_b == 0) {
    if (nnz_c_host) *nnz_c_host = 0;
    return 0;
  }
  if (!rowptr_a || !col_a || !rowptr_b || !col_b) return TSB200_ERR_INVALID_ARG;
  SpParams p;
  p.rowptr_a = rowptr_a; p.col_a = col_a; p.val_a = nullptr;
  p.rowptr_b = rowptr_b; p.col_b = col_b; p.val_b = nullptr;
  p.M = M; p.Kd = Kd; p.N = N;
  p.counts = rowptr_c + 1; p.rowptr_c = nullptr; p.row_c = nullptr; p.col_c = nullptr; p.val_c = nullptr;
  p.counter = (unsigned int*)(ws + L.scalars);
  p.window_bits = window_bits_for(N);
  p.acc_cap = 0;
  int rc = sp_launch<false, float>(p, st);
  if (rc) return rc;
  size_t tb = L.cub_bytes;
  TSB_CUDA_TRY(cub::DeviceScan::InclusiveSum(ws + L.cub, tb, rowptr_c + 1, rowptr_c + 1, (int)M, st));
  int64_t* nnz_dev = (int64_t*)(ws + L.scalars + 16);
  sp_copy_last_kernel<<<1, 1, 0, st>>>(rowptr_c, M, nnz_dev);
  TSB_LAUNCH_CHECK();
  if (nnz_c_host) TSB_CUDA_TRY(cudaMemcpyAsync(nnz_c_host, nnz_dev, 8, cudaMemcpyDeviceToHost, st));
  return 0;
}

extern "C" int tsb200_spspmm_numeric(const int64_t* rowptr_a, const int64_t* col_a, const void* val_a,
                                     const int64_t* rowptr_b, const int64_t* col_b, const void* val_b, int64_t M,
                                     int64_t Kd, int64_t N, int64_t nnz_a, int64_t nnz_b, const int64_t* rowptr_c,
                                     int64_t* row_c, int64_t* col_c, void* val_c, int dtype, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (M < 0 || Kd < 0 || N < 0 || nnz_a < 0 || nnz_b < 0) return TSB200_ERR_INVALID_ARG;
  if (M == 0 || nnz_a == 0 || nnz_b == 0) return 0;
  if (!rowptr_a || !col_a || !rowptr_b || !col_b || !rowptr_c || !col_c) return TSB200_ERR_INVALID_ARG;
  if (val_c && dtype != TSB200_F32 && dtype != TSB200_F64) return TSB200_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const SpLayout L = sp_layout(M);
  if (!workspace || workspace_bytes < L.total) return TSB200_ERR_WORKSPACE;
  char* ws = (char*)workspace;
  TSB_CUDA_TRY(cudaMemsetAsync(ws + L.scalars, 0, 256, st));
  SpParams p;
  p.rowptr_a = rowptr_a; p.col_a = col_a; p.val_a = val_a;
  p.rowptr_b = rowptr_b; p.col_b = col_b; p.val_b = val_b;
  p.M = M; p.Kd = Kd; p.N = N;
  p.counts = nullptr; p.rowptr_c = rowptr_c; p.row_c = row_c; p.col_c = col_c; p.val_c = val_c;
  p.counter = (unsigned int*)(ws + L.scalars);
  p.window_bits = window_bits_for(N);
  {
    // 0 = accumulate with L2 atomics straight into val_c (measured faster than a 16 KB shared accumulator:
    // 4 instead of 3 resident CTAs/SM and two fewer passes); the knob stays for experiments
    static const int acc = getenv("TSB200_SPSPMM_ACC") ? atoi(getenv("TSB200_SPSPMM_ACC")) : 0;
    p.acc_cap = val_c ? acc : 0;
  }
  if (val_c && dtype == TSB200_F64) return sp_launch<true, double>(p, st);
  return sp_launch<true, float>(p, st);
}
