// Pruning score -> exact global k-th smallest -> mask, for sm_100a.
//
// Replaces utils/pruning_utils.py:73-87 / :186-203 / :263-283 of the reference
// (per-layer score temporaries, torch.cat, single-CTA 16-pass torch.kthvalue, per-layer
// torch.where) with a bracketed single sweep over the weights:
//
//   1. k_sample_coarse: 2^20 strided samples -> sample keys + 2048-bin histogram of key bits [30:20]
//   2. k_sample_fine  : histograms of bits [19:9] inside the coarse bins that hold the sample
//                       quantile +- 5 sigma of its rank error (shared-memory privatised, no hot atomics);
//                       the sweep's prologue turns them into a key bracket [lo, hi) that contains
//                       the true k-th key w.h.p.
//   3. k_sweep       : ONE pass over (w, m[, g]) at HBM rate: counts keys < lo and == lo,
//                      writes the final mask for every key outside (lo, hi) and appends the
//                      few keys inside to a candidate list            (12 B/elem mag, 16 snip)
//   4. k_resolve     : one CTA selects the exact (k - n_lt - n_eq)-th candidate (8-bit radix,
//                      data in L2), emits the threshold and patches the candidates' masks.
//
// If the bracket misses (adversarial ties, overflow of the candidate list) the host falls
// back to an exact 3-pass 11/11/10-bit radix select over the full data + one apply pass.
// Both paths are bit-exact with torch.kthvalue + torch.where(score <= thr, 0, 1).
//
// Keys: scores are |.| of fp32 products, so their IEEE bit patterns order like unsigned
// integers; NaN patterns (> 0x7f800000) sort above +inf exactly like ATen's radix key
// (SortingRadixSelect.cuh:20-39).
#include "tp_common.cuh"

namespace tp {

struct SelState {
  unsigned long long k;
  unsigned long long n_lt, n_eq, n_cand;
  unsigned int lo, hi;           // bracket: lo inclusive lower edge, hi exclusive upper edge
  unsigned int thr_key;
  int status;                    // 0 ok, 1 bracket missed -> fallback, 2 ok but threshold is NaN
  unsigned int prefix, prefix_mask;   // exact radix path
  unsigned int c_lo, c_hi;            // coarse sample bins of the two bracket ranks
  unsigned long long before_lo, before_hi;
  unsigned long long k_rem;
};

constexpr int kSampleBits = 20;
constexpr int kSweepThreads = 256;
constexpr int kSmemCand = 1024;

template <int KIND>
__device__ __forceinline__ unsigned int score_key(float w, float g, float m) {
  float s;
  if (KIND == TP_SCORE_MAG) s = m * w;                 // pruning_utils.py:75
  else if (KIND == TP_SCORE_SNIP) s = (g * w) * m;     // pruning_utils.py:190
  else s = (m * g) * w;                                // pruning_utils.py:267
  return __float_as_uint(fabsf(s));
}

template <int KIND>
__device__ __forceinline__ unsigned int seg_key(const Seg& sg, long long i) {
  float w = sg.w[i], m = sg.m[i];
  float g = (KIND == TP_SCORE_MAG) ? 0.f : sg.g[i];
  return score_key<KIND>(w, g, m);
}

// ---------------------------------------------------------------------------------------------
// Block-wide rank search in a histogram of blockDim.x * PER bins: smallest bin whose inclusive
// cumulative count reaches `rank` (1-indexed).  All threads call; result through shared memory.
template <int PER>
__device__ __forceinline__ void block_find_rank(const unsigned int* __restrict__ hist, unsigned long long rank,
                                                unsigned int* s_bin, unsigned long long* s_before,
                                                unsigned long long* s_warp /* [32] */) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
  unsigned int h[PER];
  unsigned long long loc = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) { h[i] = hist[t * PER + i]; loc += h[i]; }
  unsigned long long inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  __syncthreads();                       // s_warp may still be read from a previous call
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned long long v = lane < nwarps ? s_warp[lane] : 0ull, x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long u = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += u;
    }
    s_warp[lane] = x - v;                // exclusive prefix of the warp totals
  }
  __syncthreads();
  const unsigned long long before = s_warp[warp] + inc - loc;
  if (rank > before && rank <= before + loc) {
    unsigned long long c = before;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if (c + h[i] >= rank) { *s_bin = (unsigned int)(t * PER + i); *s_before = c; break; }
      c += h[i];
    }
  }
  __syncthreads();
}

// Bracket search, level 1: strided sample -> keys (kept for level 2) + 2048-bin histogram of key
// bits [30:20], privatised in shared memory (no hot global atomics).
constexpr int kCoarseShift = 20, kFineShift = 9, kDigitBins = 2048;

template <int KIND>
__global__ void __launch_bounds__(256) k_sample_coarse(const Seg* __restrict__ segs, int n_seg, long long N, long long S,
                                                       unsigned int* __restrict__ skeys, unsigned int* __restrict__ hist_c) {
  __shared__ unsigned int s_h[kDigitBins];
  const int t = threadIdx.x;
  for (int i = t; i < kDigitBins; i += 256) s_h[i] = 0;
  __syncthreads();
  // samples are taken in runs of 4 neighbouring elements: one 32-byte sector per operand serves 4 samples
  // (a quarter of the scattered DRAM traffic and of the segment searches of single-element sampling)
  const long long G = (S + 3) >> 2;
  for (long long gi = blockIdx.x * 256ll + t; gi < G; gi += (long long)gridDim.x * 256) {
    long long e = (long long)(((unsigned long long)gi * (unsigned long long)N) / (unsigned long long)G);
    const int si = find_seg_by_elem(segs, n_seg, e);
    const Seg& sg = segs[si];
    long long l0 = (e - sg.start) & ~3ll;
    if (l0 + 3 >= sg.n) l0 = sg.n >= 4 ? sg.n - 4 : 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long j = gi * 4 + u;
      if (j >= S) break;
      const long long l = l0 + u < sg.n ? l0 + u : sg.n - 1;
      const unsigned int key = seg_key<KIND>(sg, l);
      skeys[j] = key;
      atomicAdd(&s_h[key >> kCoarseShift], 1u);
    }
  }
  __syncthreads();
  for (int i = t; i < kDigitBins; i += 256) if (s_h[i]) atomicAdd(&hist_c[i], s_h[i]);
}

// Level 2: every CTA locates the coarse bins of the two bracket ranks (8 KB of L2 reads), then the
// sample keys inside those bins are histogrammed on bits [19:9].
__global__ void __launch_bounds__(256) k_sample_fine(const unsigned int* __restrict__ skeys, long long S,
                                                     const unsigned int* __restrict__ hist_c,
                                                     long long r_lo, long long r_hi,
                                                     unsigned int* __restrict__ hist_f, SelState* st) {
  __shared__ unsigned int s_h[2][kDigitBins];
  __shared__ unsigned int s_bin[2];
  __shared__ unsigned long long s_before[2], s_warp[32];
  const int t = threadIdx.x;
  if (t < 2) { s_bin[t] = 0; s_before[t] = 0; }
  for (int i = t; i < 2 * kDigitBins; i += 256) (&s_h[0][0])[i] = 0;
  __syncthreads();
  block_find_rank<kDigitBins / 256>(hist_c, (unsigned long long)(r_lo < 1 ? 1 : r_lo), &s_bin[0], &s_before[0], s_warp);
  block_find_rank<kDigitBins / 256>(hist_c, (unsigned long long)(r_hi > S ? S : r_hi), &s_bin[1], &s_before[1], s_warp);
  const unsigned int c_lo = s_bin[0], c_hi = s_bin[1];
  if (blockIdx.x == 0 && t == 0) { st->c_lo = c_lo; st->c_hi = c_hi; st->before_lo = s_before[0]; st->before_hi = s_before[1]; }
  for (long long j = blockIdx.x * 256ll + t; j < S; j += (long long)gridDim.x * 256) {
    const unsigned int key = skeys[j], c = key >> kCoarseShift, f = (key >> kFineShift) & (kDigitBins - 1);
    if (c == c_lo) atomicAdd(&s_h[0][f], 1u);
    if (c == c_hi) atomicAdd(&s_h[1][f], 1u);
  }
  __syncthreads();
  for (int i = t; i < 2 * kDigitBins; i += 256) { unsigned int v = (&s_h[0][0])[i]; if (v) atomicAdd(&hist_f[i], v); }
}

// ---------------------------------------------------------------------------------------------
struct SweepCtx {
  unsigned int lo, hi;
  unsigned int n_lt, n_eq;
  uint2* s_cand; unsigned int* s_ncand;
  SelState* st; uint2* cand; unsigned int cap;
};

__device__ __forceinline__ float classify(SweepCtx& c, unsigned int key, long long gidx) {
  if (key < c.lo) { c.n_lt++; return 0.f; }
  if (key == c.lo) { c.n_eq++; return 0.f; }
  if (key >= c.hi) return 1.f;
  // inside the bracket: rare (~0.1-1 %) -> candidate list
  unsigned int slot = atomicAdd(c.s_ncand, 1u);
  if (slot < kSmemCand) {
    c.s_cand[slot] = make_uint2(key, (unsigned int)gidx);
  } else {
    unsigned long long gs = atomicAdd(&c.st->n_cand, 1ull);
    if (gs < c.cap) c.cand[gs] = make_uint2(key, (unsigned int)gidx);
  }
  return 0.f;   // provisional; k_resolve patches it
}

template <int KIND, bool WRITE>
__global__ void __launch_bounds__(kSweepThreads) k_sweep(const Seg* __restrict__ segs, int n_seg, long long tiles,
                                                         SelState* st, uint2* __restrict__ cand, unsigned int cap,
                                                         const unsigned int* __restrict__ hist_f,
                                                         long long r_lo, long long r_hi, long long S) {
  __shared__ uint2 s_cand[kSmemCand];
  __shared__ unsigned int s_ncand;
  __shared__ unsigned long long s_base;
  __shared__ unsigned int s_red[2][kSweepThreads / 32];
  __shared__ unsigned int s_bin[2];
  __shared__ unsigned long long s_before[2], s_warp[32];
  // bracket [lo, hi) from the fine sample histograms — every CTA derives the same two keys
  if (threadIdx.x < 2) { s_bin[threadIdx.x] = 0; s_before[threadIdx.x] = 0; }
  __syncthreads();
  block_find_rank<kDigitBins / kSweepThreads>(hist_f, (unsigned long long)(r_lo < 1 ? 1 : r_lo) - st->before_lo,
                                              &s_bin[0], &s_before[0], s_warp);
  block_find_rank<kDigitBins / kSweepThreads>(hist_f + kDigitBins, (unsigned long long)(r_hi > S ? S : r_hi) - st->before_hi,
                                              &s_bin[1], &s_before[1], s_warp);
  unsigned int lo = (st->c_lo << kCoarseShift) | (s_bin[0] << kFineShift);
  unsigned int hi = (((st->c_hi << 11) | s_bin[1]) + 1u) << kFineShift;
  if (r_lo < 1) lo = 0u;                              // no lower bound
  if (r_hi > S || hi > 0x80000000u || hi == 0u) hi = 0x80000000u;   // no upper bound (keys are <= 0x7fffffff)
  if (blockIdx.x == 0 && threadIdx.x == 0) { st->lo = lo; st->hi = hi; }
  SweepCtx c;
  c.lo = lo; c.hi = hi; c.n_lt = 0; c.n_eq = 0;
  c.s_cand = s_cand; c.s_ncand = &s_ncand; c.st = st; c.cand = cand; c.cap = cap;
  const int t = threadIdx.x;
  constexpr int kVecIters = kTileElems / (kSweepThreads * 4);   // 4

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    if (t == 0) s_ncand = 0;
    __syncthreads();
    const int si = find_seg(segs, n_seg, tile);
    const Seg sg = segs[si];
    const long long base = (tile - sg.tile0) * kTileElems;
    const long long rem = sg.n - base;
    const int n_in = rem < kTileElems ? (int)rem : kTileElems;
    const float* wp = sg.w + base;
    const float* mp = sg.m + base;
    const float* gp = (KIND == TP_SCORE_MAG) ? nullptr : sg.g + base;
    float* op = WRITE ? sg.mo + base : nullptr;
    const long long g0 = sg.start + base;
    bool vec = n_in == kTileElems &&
               ((((uintptr_t)wp) | ((uintptr_t)mp) | ((uintptr_t)gp) | ((uintptr_t)op)) & 15) == 0;
    if (vec) {
      float4 wv[kVecIters], mv[kVecIters], gv[kVecIters];
#pragma unroll
      for (int it = 0; it < kVecIters; ++it) {
        int q = it * kSweepThreads + t;
        wv[it] = ld_stream((const float4*)wp + q);
        mv[it] = ld_stream((const float4*)mp + q);
        if (KIND != TP_SCORE_MAG) gv[it] = ld_stream((const float4*)gp + q);
        else gv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < kVecIters; ++it) {
        int q = it * kSweepThreads + t;
        long long gi = g0 + (long long)q * 4;
        float4 o;
        o.x = classify(c, score_key<KIND>(wv[it].x, gv[it].x, mv[it].x), gi + 0);
        o.y = classify(c, score_key<KIND>(wv[it].y, gv[it].y, mv[it].y), gi + 1);
        o.z = classify(c, score_key<KIND>(wv[it].z, gv[it].z, mv[it].z), gi + 2);
        o.w = classify(c, score_key<KIND>(wv[it].w, gv[it].w, mv[it].w), gi + 3);
        if (WRITE) st_stream((float4*)op + q, o);
      }
    } else {
      for (int i = t; i < n_in; i += kSweepThreads) {
        float g = (KIND == TP_SCORE_MAG) ? 0.f : gp[i];
        float o = classify(c, score_key<KIND>(wp[i], g, mp[i]), g0 + i);
        if (WRITE) op[i] = o;
      }
    }
    __syncthreads();
    unsigned int nc = s_ncand < (unsigned)kSmemCand ? s_ncand : (unsigned)kSmemCand;
    if (nc) {
      if (t == 0) s_base = atomicAdd(&st->n_cand, (unsigned long long)nc);
      __syncthreads();
      unsigned long long b = s_base;
      for (unsigned int i = t; i < nc; i += kSweepThreads)
        if (b + i < cap) cand[b + i] = s_cand[i];
    }
    __syncthreads();
  }
  // block-reduce the two counters, one atomic pair per CTA
  unsigned int a = c.n_lt, b = c.n_eq;
  for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((t & 31) == 0) { s_red[0][t >> 5] = a; s_red[1][t >> 5] = b; }
  __syncthreads();
  if (t == 0) {
    unsigned long long A = 0, B = 0;
    for (int i = 0; i < kSweepThreads / 32; ++i) { A += s_red[0][i]; B += s_red[1][i]; }
    if (A) atomicAdd(&st->n_lt, A);
    if (B) atomicAdd(&st->n_eq, B);
  }
}

// ---------------------------------------------------------------------------------------------
// Resolve, stage 1 (multi-CTA): decide success of the bracket and histogram the FIRST radix digit of all
// candidates (the pass that touches every candidate) with shared-memory privatised bins.
__device__ __forceinline__ int resolve_shift0(unsigned int lo, unsigned int hi) {
  const unsigned int span = lo ^ (hi - 1u);
  const int top = span ? (31 - __clz(span)) : 0;
  return (top / 11) * 11;
}

__global__ void __launch_bounds__(256) k_cand_hist(const SelState* __restrict__ st, const uint2* __restrict__ cand,
                                                   unsigned int cap, unsigned int* __restrict__ hist_r) {
  __shared__ unsigned int s_h[kDigitBins];
  const unsigned long long k = st->k, A = st->n_lt, B = st->n_eq, C = st->n_cand;
  if (k <= A + B || C > cap || k > A + B + C) return;        // nothing to select (k_resolve decides what it means)
  const int t = threadIdx.x;
  for (int i = t; i < kDigitBins; i += 256) s_h[i] = 0;
  __syncthreads();
  const int shift = resolve_shift0(st->lo, st->hi);
  const unsigned int n = (unsigned int)C;
  for (unsigned int i = blockIdx.x * 256u + t; i < n; i += gridDim.x * 256u)
    atomicAdd(&s_h[(cand[i].x >> shift) & (kDigitBins - 1)], 1u);
  __syncthreads();
  for (int i = t; i < kDigitBins; i += 256) if (s_h[i]) atomicAdd(&hist_r[i], s_h[i]);
}

// Resolve, stage 2 (one CTA): pick the first digit from the global histogram, finish the remaining digits over
// the (now ~1/2048-th) matching candidates, publish threshold and status.
__global__ void __launch_bounds__(1024) k_resolve(SelState* st, const uint2* __restrict__ cand, unsigned int cap,
                                                  const unsigned int* __restrict__ hist_r, float* __restrict__ thr_out) {
  __shared__ unsigned int s_hist[kDigitBins];
  __shared__ unsigned int s_prefix, s_bin;
  __shared__ unsigned long long s_before, s_krem, s_warp[32];
  __shared__ int s_status;
  const int t = threadIdx.x;
  const unsigned long long k = st->k, A = st->n_lt, B = st->n_eq, C = st->n_cand;
  if (t == 0) {
    int status = 0;
    if (k <= A || C > cap) status = 1;            // k-th lies below the bracket / list overflow
    else if (k <= A + B) { st->thr_key = st->lo; }
    else if (k > A + B + C) status = 1;           // k-th lies above the bracket
    s_status = status;
    s_krem = k - A - B;
  }
  __syncthreads();
  if (s_status == 1) { if (t == 0) st->status = 1; return; }
  const unsigned int n = (unsigned int)C;
  constexpr int U = 8;                            // independent L2 loads in flight per thread
  if (k > A + B) {
    const int shift0 = resolve_shift0(st->lo, st->hi);
    if (t == 0) { s_prefix = (shift0 + 11 >= 32) ? 0u : (st->lo & (0xFFFFFFFFu << (shift0 + 11))); s_bin = 0; s_before = 0; }
    __syncthreads();
    block_find_rank<kDigitBins / 1024>(hist_r, s_krem, &s_bin, &s_before, s_warp);
    if (t == 0) { s_prefix |= (s_bin << shift0); s_krem -= s_before; }
    __syncthreads();
    for (int shift = shift0 - 11; shift >= 0; shift -= 11) {
      for (int i = t; i < kDigitBins; i += 1024) s_hist[i] = 0;
      if (t == 0) { s_bin = 0; s_before = 0; }
      __syncthreads();
      const unsigned int prefix = s_prefix;
      const unsigned int pmask = 0xFFFFFFFFu << (shift + 11);
      for (unsigned int base = 0; base < n; base += 1024 * U) {
        unsigned int key[U]; bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned int i = base + u * 1024 + t;
          ok[u] = i < n;
          key[u] = ok[u] ? cand[i].x : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (ok[u] && (key[u] & pmask) == prefix) atomicAdd(&s_hist[(key[u] >> shift) & (kDigitBins - 1)], 1u);
      }
      __syncthreads();
      block_find_rank<kDigitBins / 1024>(s_hist, s_krem, &s_bin, &s_before, s_warp);
      if (t == 0) { s_prefix = prefix | (s_bin << shift); s_krem -= s_before; }
      __syncthreads();
    }
    if (t == 0) st->thr_key = s_prefix;
  }
  __syncthreads();
  if (t == 0) {
    const unsigned int thr = st->thr_key;
    *thr_out = __uint_as_float(thr);
    st->status = (thr > 0x7f800000u) ? 2 : 0;
  }
}

// Resolve, stage 3 (multi-CTA): final mask value of every candidate (the sweep wrote a provisional 0).
__global__ void __launch_bounds__(256) k_cand_patch(const Seg* __restrict__ segs, int n_seg, const SelState* __restrict__ st,
                                                    const uint2* __restrict__ cand, unsigned int cap) {
  if (st->status != 0 || st->n_cand > cap) return;          // fallback / NaN threshold: another kernel rewrites every mask
  const unsigned int thr = st->thr_key, n = (unsigned int)st->n_cand;
  for (unsigned int i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) {
    const uint2 c = cand[i];
    const int si = find_seg_by_elem(segs, n_seg, (long long)c.y);
    segs[si].mo[(long long)c.y - segs[si].start] = (c.x <= thr) ? 0.f : 1.f;
  }
}

// ---------------------------------------------------------------------------------------------
// Exact fallback: 11/11/10-bit radix passes over the full data.
template <int KIND>
__global__ void __launch_bounds__(kSweepThreads) k_hist_pass(const Seg* __restrict__ segs, int n_seg, long long tiles,
                                                             const SelState* __restrict__ st, int shift, int nbits,
                                                             unsigned int* __restrict__ hist) {
  __shared__ unsigned int s_hist[2048];
  const int t = threadIdx.x;
  for (int i = t; i < 2048; i += kSweepThreads) s_hist[i] = 0;
  __syncthreads();
  const unsigned int prefix = st->prefix, pmask = st->prefix_mask, dmask = (1u << nbits) - 1u;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int si = find_seg(segs, n_seg, tile);
    const Seg& sg = segs[si];
    const long long base = (tile - sg.tile0) * kTileElems;
    const long long rem = sg.n - base;
    const int n_in = rem < kTileElems ? (int)rem : kTileElems;
    for (int i = t; i < n_in; i += kSweepThreads) {
      unsigned int key = seg_key<KIND>(sg, base + i);
      if ((key & pmask) == prefix) atomicAdd(&s_hist[(key >> shift) & dmask], 1u);
    }
  }
  __syncthreads();
  for (int i = t; i < (1 << nbits); i += kSweepThreads)
    if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

__global__ void k_pick_digit(const unsigned int* __restrict__ hist, int shift, int nbits, SelState* st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long cum = 0, kr = st->k_rem;
  for (int d = 0; d < (1 << nbits); ++d) {
    unsigned long long h = hist[d];
    if (cum + h >= kr) {
      st->prefix |= ((unsigned int)d << shift);
      st->prefix_mask |= (((1u << nbits) - 1u) << shift);
      st->k_rem = kr - cum;
      break;
    }
    cum += h;
  }
}

__global__ void k_publish_thr(SelState* st, float* thr_out) {
  st->thr_key = st->prefix;
  st->status = (st->prefix > 0x7f800000u) ? 2 : 0;
  *thr_out = __uint_as_float(st->prefix);
}

// mask_out = score <= thr ? 0 : 1 in fp32 compare semantics (NaN threshold keeps everything).
template <int KIND>
__global__ void __launch_bounds__(kSweepThreads) k_apply(const Seg* __restrict__ segs, int n_seg, long long tiles,
                                                         const float* __restrict__ thr_p) {
  const float thr = *thr_p;
  const int t = threadIdx.x;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int si = find_seg(segs, n_seg, tile);
    const Seg sg = segs[si];
    const long long base = (tile - sg.tile0) * kTileElems;
    const long long rem = sg.n - base;
    const int n_in = rem < kTileElems ? (int)rem : kTileElems;
    const float* wp = sg.w + base;
    const float* mp = sg.m + base;
    const float* gp = (KIND == TP_SCORE_MAG) ? nullptr : sg.g + base;
    float* op = sg.mo + base;
    bool vec = n_in == kTileElems &&
               ((((uintptr_t)wp) | ((uintptr_t)mp) | ((uintptr_t)gp) | ((uintptr_t)op)) & 15) == 0;
    if (vec) {
#pragma unroll
      for (int it = 0; it < kTileElems / (kSweepThreads * 4); ++it) {
        int q = it * kSweepThreads + t;
        float4 w = ld_stream((const float4*)wp + q), m = ld_stream((const float4*)mp + q);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (KIND != TP_SCORE_MAG) g = ld_stream((const float4*)gp + q);
        float4 o;
        o.x = (__uint_as_float(score_key<KIND>(w.x, g.x, m.x)) <= thr) ? 0.f : 1.f;
        o.y = (__uint_as_float(score_key<KIND>(w.y, g.y, m.y)) <= thr) ? 0.f : 1.f;
        o.z = (__uint_as_float(score_key<KIND>(w.z, g.z, m.z)) <= thr) ? 0.f : 1.f;
        o.w = (__uint_as_float(score_key<KIND>(w.w, g.w, m.w)) <= thr) ? 0.f : 1.f;
        st_stream((float4*)op + q, o);
      }
    } else {
      for (int i = t; i < n_in; i += kSweepThreads) {
        float g = (KIND == TP_SCORE_MAG) ? 0.f : gp[i];
        op[i] = (__uint_as_float(score_key<KIND>(wp[i], g, mp[i])) <= thr) ? 0.f : 1.f;
      }
    }
  }
}

__global__ void __launch_bounds__(kSweepThreads) k_count_zeros(const Seg* __restrict__ segs, int n_seg, long long tiles,
                                                               unsigned long long* __restrict__ out) {
  __shared__ unsigned int s_red[kSweepThreads / 32];
  const int t = threadIdx.x;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int si = find_seg(segs, n_seg, tile);
    const Seg sg = segs[si];
    const long long base = (tile - sg.tile0) * kTileElems;
    const long long rem = sg.n - base;
    const int n_in = rem < kTileElems ? (int)rem : kTileElems;
    const float* mp = sg.m + base;
    unsigned int z = 0;
    if (n_in == kTileElems && (((uintptr_t)mp) & 15) == 0) {
#pragma unroll
      for (int it = 0; it < kTileElems / (kSweepThreads * 4); ++it) {
        float4 m = ld_stream((const float4*)mp + it * kSweepThreads + t);
        z += (m.x == 0.f) + (m.y == 0.f) + (m.z == 0.f) + (m.w == 0.f);
      }
    } else {
      for (int i = t; i < n_in; i += kSweepThreads) z += (mp[i] == 0.f);
    }
    for (int o = 16; o; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    if ((t & 31) == 0) s_red[t >> 5] = z;
    __syncthreads();
    if (t == 0) {
      unsigned long long tot = 0;
      for (int i = 0; i < kSweepThreads / 32; ++i) tot += s_red[i];
      if (tot) { atomicAdd(&out[si], tot); atomicAdd(&out[n_seg], tot); }
    }
    __syncthreads();
  }
}

static unsigned int cand_cap(long long N) {
  long long c = N / 8;
  if (c < (1 << 16)) c = 1 << 16;
  if (c > (1 << 23)) c = 1 << 23;
  return (unsigned int)c;
}

static int sweep_grid(long long tiles) {
  long long g = (long long)sm_count() * 8;
  return (int)(tiles < g ? (tiles > 0 ? tiles : 1) : g);
}

template <int KIND>
static int run_topk(const Seg* d_segs, int n_seg, long long tiles, long long N, long long k, bool write,
                    SelState* d_st, unsigned int* d_hist, unsigned int* d_skeys, uint2* d_cand, unsigned int cap,
                    float* thr_out, int64_t* info, cudaStream_t st) {
  SelState h = {};
  h.k = (unsigned long long)k;
  h.hi = 0x80000000u;
  TP_CUDA_CHECK(cudaMemcpyAsync(d_st, &h, sizeof(h), cudaMemcpyHostToDevice, st));
  // d_hist layout: [0,2048) coarse sample histogram, [2048, 6144) two fine histograms, [6144, 8192) candidate digit
  TP_CUDA_CHECK(cudaMemsetAsync(d_hist, 0, sizeof(unsigned int) * 4 * kDigitBins, st));
  const long long S = N < (1ll << kSampleBits) ? N : (1ll << kSampleBits);
  // sample rank of the population's k-th element and the +-5 sigma band of its sampling error
  long long rs = (long long)(((unsigned __int128)(unsigned long long)k * (unsigned long long)S + (unsigned long long)N - 1) /
                             (unsigned long long)N);
  if (rs < 1) rs = 1;
  if (rs > S) rs = S;
  const double pq = (double)rs / (double)S;
  // +4 per segment: runs are clamped at segment ends, so a few samples may repeat (also when S == N)
  const long long delta = (long long)(5.0 * sqrt((double)S * pq * (1.0 - pq)) + 8.0) + 4ll * n_seg;
  const long long r_lo = rs - delta, r_hi = rs + delta;
  const int sgrid = (int)((S + 255) / 256 < (long long)sm_count() * 4 ? (S + 255) / 256 : (long long)sm_count() * 4);
  k_sample_coarse<KIND><<<sgrid, 256, 0, st>>>(d_segs, n_seg, N, S, d_skeys, d_hist);
  k_sample_fine<<<sgrid, 256, 0, st>>>(d_skeys, S, d_hist, r_lo, r_hi, d_hist + kDigitBins, d_st);
  const int grid = sweep_grid(tiles);
  if (write) k_sweep<KIND, true><<<grid, kSweepThreads, 0, st>>>(d_segs, n_seg, tiles, d_st, d_cand, cap, d_hist + kDigitBins, r_lo, r_hi, S);
  else       k_sweep<KIND, false><<<grid, kSweepThreads, 0, st>>>(d_segs, n_seg, tiles, d_st, d_cand, cap, d_hist + kDigitBins, r_lo, r_hi, S);
  k_cand_hist<<<64, 256, 0, st>>>(d_st, d_cand, cap, d_hist + 3 * kDigitBins);
  k_resolve<<<1, 1024, 0, st>>>(d_st, d_cand, cap, d_hist + 3 * kDigitBins, thr_out);
  if (write) k_cand_patch<<<64, 256, 0, st>>>(d_segs, n_seg, d_st, d_cand, cap);
  TP_LAUNCH_CHECK();
  TP_CUDA_CHECK(cudaMemcpyAsync(&h, d_st, sizeof(h), cudaMemcpyDeviceToHost, st));
  TP_CUDA_CHECK(cudaStreamSynchronize(st));
  int path = 0;
  if (h.status == 1) {
    // exact fallback: 3 radix passes over the full data
    path = 1;
    SelState f = {};
    f.k = (unsigned long long)k; f.k_rem = (unsigned long long)k;
    TP_CUDA_CHECK(cudaMemcpyAsync(d_st, &f, sizeof(f), cudaMemcpyHostToDevice, st));
    const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
    for (int p = 0; p < 3; ++p) {
      TP_CUDA_CHECK(cudaMemsetAsync(d_hist, 0, sizeof(unsigned int) * 2048, st));
      k_hist_pass<KIND><<<grid, kSweepThreads, 0, st>>>(d_segs, n_seg, tiles, d_st, shifts[p], bits[p], d_hist);
      k_pick_digit<<<1, 32, 0, st>>>(d_hist, shifts[p], bits[p], d_st);
    }
    k_publish_thr<<<1, 1, 0, st>>>(d_st, thr_out);
    if (write) k_apply<KIND><<<grid, kSweepThreads, 0, st>>>(d_segs, n_seg, tiles, thr_out);
    TP_LAUNCH_CHECK();
    TP_CUDA_CHECK(cudaMemcpyAsync(&h, d_st, sizeof(h), cudaMemcpyDeviceToHost, st));
    TP_CUDA_CHECK(cudaStreamSynchronize(st));
  } else if (h.status == 2 && write) {
    // NaN threshold: `score <= nan` is false everywhere -> every mask entry becomes 1
    k_apply<KIND><<<grid, kSweepThreads, 0, st>>>(d_segs, n_seg, tiles, thr_out);
    TP_LAUNCH_CHECK();
  }
  if (info) { info[0] = path; info[1] = (int64_t)h.n_cand; info[2] = (int64_t)h.n_lt; info[3] = (h.status == 2); }
  return TP_OK;
}

}  // namespace tp

using namespace tp;

extern "C" {

size_t tp_topk_workspace_bytes(int n_seg, int64_t total_numel) {
  size_t b = 0;
  b += align_up(sizeof(Seg) * (size_t)(n_seg > 0 ? n_seg : 1), 256);
  b += align_up(sizeof(SelState), 256);
  b += align_up(sizeof(unsigned int) * 4 * kDigitBins, 256);
  b += align_up(sizeof(unsigned int) * (size_t)(1u << kSampleBits), 256);
  b += align_up(sizeof(uint2) * (size_t)cand_cap(total_numel), 256);
  return b + 1024;
}

int tp_topk_threshold_mask(const void* const* w, const void* const* g, const void* const* m,
                           void* const* mask_out, const int64_t* numel, int n_seg,
                           int64_t k, int score_kind, float* thr_out,
                           void* ws, size_t ws_bytes, int64_t* info_out, void* stream) {
  if (!w || !m || !numel || n_seg <= 0 || !thr_out || !ws) return TP_ERR_INVALID;
  if (score_kind != TP_SCORE_MAG && !g) return TP_ERR_INVALID;
  if (score_kind < 0 || score_kind > 2) return TP_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  long long N = 0;
  for (int i = 0; i < n_seg; ++i) { if (numel[i] < 0) return TP_ERR_INVALID; N += numel[i]; }
  if (k < 1 || k > N) return TP_ERR_K_RANGE;          // torch.kthvalue raises (pruning_utils.py:79)
  if (N >= (1ll << 32)) return TP_ERR_UNSUPPORTED;    // candidate records carry 32-bit indices
  Arena ar(ws, ws_bytes);
  Seg* d_segs = nullptr; long long tiles = 0, total = 0;
  int rc = upload_segs(ar, w, g, m, mask_out, nullptr, numel, n_seg, &d_segs, &tiles, &total, st);
  if (rc) return rc;
  SelState* d_st = (SelState*)ar.take(sizeof(SelState));
  unsigned int* d_hist = (unsigned int*)ar.take(sizeof(unsigned int) * 4 * kDigitBins);
  unsigned int* d_skeys = (unsigned int*)ar.take(sizeof(unsigned int) * (size_t)(1u << kSampleBits));
  const unsigned int cap = cand_cap(N);
  uint2* d_cand = (uint2*)ar.take(sizeof(uint2) * (size_t)cap);
  if (!d_st || !d_hist || !d_skeys || !d_cand) return TP_ERR_WORKSPACE;
  const bool write = mask_out != nullptr;
  switch (score_kind) {
    case TP_SCORE_MAG: return run_topk<TP_SCORE_MAG>(d_segs, n_seg, tiles, N, k, write, d_st, d_hist, d_skeys, d_cand, cap, thr_out, info_out, st);
    case TP_SCORE_SNIP: return run_topk<TP_SCORE_SNIP>(d_segs, n_seg, tiles, N, k, write, d_st, d_hist, d_skeys, d_cand, cap, thr_out, info_out, st);
    default: return run_topk<TP_SCORE_SYNFLOW>(d_segs, n_seg, tiles, N, k, write, d_st, d_hist, d_skeys, d_cand, cap, thr_out, info_out, st);
  }
}

int tp_apply_threshold(const void* const* w, const void* const* g, const void* const* m,
                       void* const* mask_out, const int64_t* numel, int n_seg,
                       int score_kind, const float* thr, void* ws, size_t ws_bytes, void* stream) {
  if (!w || !m || !mask_out || !numel || n_seg <= 0 || !thr || !ws) return TP_ERR_INVALID;
  if (score_kind != TP_SCORE_MAG && !g) return TP_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  Seg* d_segs = nullptr; long long tiles = 0;
  int rc = upload_segs(ar, w, g, m, mask_out, nullptr, numel, n_seg, &d_segs, &tiles, nullptr, st);
  if (rc) return rc;
  if (tiles == 0) return TP_OK;
  const int grid = sweep_grid(tiles);
  if (score_kind == TP_SCORE_MAG) k_apply<TP_SCORE_MAG><<<grid, kSweepThreads, 0, st>>>(d_segs, n_seg, tiles, thr);
  else if (score_kind == TP_SCORE_SNIP) k_apply<TP_SCORE_SNIP><<<grid, kSweepThreads, 0, st>>>(d_segs, n_seg, tiles, thr);
  else if (score_kind == TP_SCORE_SYNFLOW) k_apply<TP_SCORE_SYNFLOW><<<grid, kSweepThreads, 0, st>>>(d_segs, n_seg, tiles, thr);
  else return TP_ERR_INVALID;
  TP_LAUNCH_CHECK();
  return TP_OK;
}

int tp_count_zeros(const void* const* m, const int64_t* numel, int n_seg,
                   int64_t* zeros_out, void* ws, size_t ws_bytes, void* stream) {
  if (!m || !numel || n_seg <= 0 || !zeros_out || !ws) return TP_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  Arena ar(ws, ws_bytes);
  Seg* d_segs = nullptr; long long tiles = 0;
  int rc = upload_segs(ar, nullptr, nullptr, m, nullptr, nullptr, numel, n_seg, &d_segs, &tiles, nullptr, st);
  if (rc) return rc;
  TP_CUDA_CHECK(cudaMemsetAsync(zeros_out, 0, sizeof(int64_t) * (size_t)(n_seg + 1), st));
  if (tiles == 0) return TP_OK;
  k_count_zeros<<<sweep_grid(tiles), kSweepThreads, 0, st>>>(d_segs, n_seg, tiles, (unsigned long long*)zeros_out);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

}  // extern "C"
