// Non-dominated rank (replaces dmosopt/dda.py:97-152 dda_ens).
//
// rank_i = front index = length of the longest domination chain ending at i
//        = 0 if nothing dominates i, else 1 + max{rank_j : j dominates i}.
//
// Pipeline (all on the context's stream):
//   1. per objective: radix-sort the column, give every distinct value a dense integer id
//      (order- and equality-preserving, so float64 inputs are ranked exactly with 32-bit compares);
//   2. lexicographic order of the integer vectors (LSD radix passes); in that order a point can only
//      be dominated by points before it, and identical vectors are adjacent (one "group id" each);
//      (two and three objectives, up to 1024 blocks: the segmented order of `RankSeg` below -- another linear extension
//      of the dominance order -- which lets whole tiles be skipped or answered from a per-tile staircase);
//   3. one persistent kernel evaluates the chain recurrence block by block in that order: a block of T = 128 targets
//      streams the earlier blocks through shared memory (coalesced 16-byte records, broadcast reads, one dominance
//      predicate per pair); the rank word of a record (rank + 1, 0 = not final) is its own ready flag, so a consumer
//      polls only when it catches up with the wavefront.  In-block chains and the chains entering from the predecessor
//      block are precomputed as longest-path tables (int8, breadth-first walks over 128-bit successor masks) before
//      any rank is needed, so that the serial part of a block is two packed max-plus products and one store.
//      Blocks are handed out by an atomic ticket, so a block only ever waits for blocks whose CTAs are already running
//      (no co-residency assumption, no deadlock); every wait is bounded by a watchdog that raises an error flag.
//
// Rank-0-only queries (the hypervolume's filter) take the cell-grid kernels `ndg_*` (M <= 3) or a plain block scan.
// Truncations (remove_worst: only the ranks of the kept rows matter) of three-objective sets peel the fronts they need off
// the same kind of grid instead of running the chain, when those fronts are few (`rank_by_peeling`, below).
//
// Algorithmic bytes: 8 n M read + 4 n written; pair tests <= n^2 / 2 (compare / latency bound).
#include <stdlib.h>

#include <cstdio>
#include <vector>

#include "common.cuh"

namespace {

constexpr int RANK_T = 128;  // targets per block == sources per shared-memory tile

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void col_keys_kernel(const double* __restrict__ Y, int64_t n, int M, int j, uint64_t* __restrict__ keys,
                                uint32_t* __restrict__ idx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = f64_to_ordered(Y[i * M + j]);
    idx[i] = (uint32_t)i;
  }
}

__global__ void flag_new_u64_kernel(const uint64_t* __restrict__ skeys, int64_t n, uint32_t* __restrict__ flag) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) flag[p] = (p > 0 && skeys[p] != skeys[p - 1]) ? 1u : 0u;
}

__global__ void scatter_dense_kernel(const uint32_t* __restrict__ dense, const uint32_t* __restrict__ sidx, int64_t n,
                                     uint32_t* __restrict__ R) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) R[sidx[p]] = dense[p];
}

__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ perm, int64_t n,
                                  uint32_t* __restrict__ out) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) out[p] = src[perm[p]];
}

// flag[p] = 1 iff the vector at sorted position p differs from the one at p-1
__global__ void flag_new_vec_kernel(const uint32_t* __restrict__ R, const uint32_t* __restrict__ perm, int64_t n, int M,
                                    uint32_t* __restrict__ flag) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t f = 0;
  if (p > 0) {
    uint32_t a = perm[p], b = perm[p - 1];
    for (int j = 0; j < M; ++j) f |= (R[(int64_t)j * n + a] != R[(int64_t)j * n + b]) ? 1u : 0u;
  }
  flag[p] = f;
}

// record words: [o_1 .. o_{M-1}, gid, rank+1, pad...]; padded to W = 4*ceil((M+1)/4) words
__global__ void build_records_kernel(const uint32_t* __restrict__ R, const uint32_t* __restrict__ perm,
                                     const uint32_t* __restrict__ gid, int64_t n, int64_t npad, int M, int W,
                                     uint32_t* __restrict__ rec) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npad) return;
  uint32_t* r = rec + p * W;
  if (p < n) {
    uint32_t i = perm[p];
    for (int j = 1; j < M; ++j) r[j - 1] = R[(int64_t)j * n + i];
    r[M - 1] = gid[p];
    for (int w = M; w < W; ++w) r[w] = 0u;
  } else {
    // sentinel: larger than every real id in every objective, its own group; never dominates a real point
    for (int j = 1; j < M; ++j) r[j - 1] = 0xFFFFFFFFu;
    r[M - 1] = 0xFFFFFFFFu - (uint32_t)(p - n);
    for (int w = M; w < W; ++w) r[w] = 0u;
  }
}

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t word_of(const uint4& a, int c) { return c == 0 ? a.x : c == 1 ? a.y : c == 2 ? a.z : a.w; }

// The rank word of a record (rank + 1, 0 = not yet final) is its own ready flag: a 4-byte store is atomic and carries
// no other data, so publication needs neither fences nor a separate flag, and consumers simply poll the word.
// ------------------------------------------------------------------------------------------------ segmented order (M <= 3)
// Any linear extension of the dominance order works for the chain recurrence.  Instead of the plain lexicographic
// order the points are ordered by (segment of objective 1, objective 2, ..., objective M, objective 1), a segment
// being a run of 1024 dense ids of objective 1 (at most 128 segments).  Sources in an earlier segment have a smaller objective-1 id than
// every target, and inside a segment the tiles are sorted by the first compare word, so for a target block with the
// band [blo, bhi] of that word a whole earlier tile is
//   * skipped        if its smallest word is > bhi (nothing in it can dominate anything in the block),
//   * "fast"         if its largest word is < blo (the first compare is known to pass and is dropped),
//   * tested in full otherwise (about one tile per segment) and, with the extra objective-1 compare, for the tiles that
//     may share a segment with the block.
// For three objectives a "fast" tile leaves a single condition, word_1(source) <= word_1(target); every block therefore
// also publishes its records as a staircase -- sorted by that word, with the running maximum of rank + 1 -- and a target
// resolves a fast tile with one 7-step binary search instead of 128 pair tests (two objectives: the tile maximum).
// Bands are compared on 8-bit floor-quantised words (conservative in both directions) held in shared memory.
struct RankSeg {
  const uint32_t* c1rec = nullptr;      // [npad] objective-1 id per position (0xFFFFFFFF for the padding)
  const uint32_t* seg_start = nullptr;  // [nseg + 1] first position whose segment is >= s
  const uint16_t* tile_q = nullptr;     // [nblocks] low byte = min, high byte = max of the quantised first compare word
  unsigned long long* stair = nullptr;  // [npad] per tile: entry t = (t-th smallest last compare word of the tile, low 32
                                        // bits; max (rank + 1) over the t+1 records with the smallest words, high 32 bits)
  int sshift = 0;                       // id1 >> sshift = segment
  int qshift = 0;                       // word >> qshift = 8-bit band coordinate (clamped to 255)
};
constexpr int RANK_SEG_MAXT = 1024;  // tiles whose bands fit the shared-memory cache (n <= 131072)

__global__ void seg_key_kernel(const uint32_t* __restrict__ R0, const uint32_t* __restrict__ perm, int64_t n, int sshift,
                               uint32_t* __restrict__ key) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) key[p] = R0[perm[p]] >> sshift;
}

__global__ void seg_c1_kernel(const uint32_t* __restrict__ R0, const uint32_t* __restrict__ perm, int64_t n, int64_t npad,
                              uint32_t* __restrict__ c1rec) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < npad) c1rec[p] = p < n ? R0[perm[p]] : 0xFFFFFFFFu;
}

__global__ void seg_start_kernel(const uint32_t* __restrict__ c1rec, int64_t n, int sshift, int nseg,
                                 uint32_t* __restrict__ seg_start) {
  int sgi = blockIdx.x * blockDim.x + threadIdx.x;
  if (sgi > nseg) return;
  int64_t lo = 0, hi = n;  // positions are sorted by segment
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((c1rec[mid] >> sshift) < (uint32_t)sgi) lo = mid + 1; else hi = mid;
  }
  seg_start[sgi] = (uint32_t)lo;
}

// one warp per tile: quantised min / max of the first compare word.  A tile whose first compare word is not ascending
// (it spans a segment boundary, or holds padding) is given the full band [0, 255]: it is then never skipped, never
// "fast", and the pair tests take the generic loop instead of the sorted-prefix loop.
__global__ void seg_tile_band_kernel(const uint32_t* __restrict__ rec, int nblocks, int T, int qshift,
                                     uint16_t* __restrict__ tile_q) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= nblocks) return;
  uint32_t lo = 255u, hi = 0u;
  bool sorted = true;
  for (int t = lane; t < T; t += 32) {
    const uint32_t w = rec[((int64_t)k * T + t) * 4];
    const uint32_t q = min(w >> qshift, 255u);
    lo = min(lo, q);
    hi = max(hi, q);
    if (t > 0) sorted = sorted && (rec[((int64_t)k * T + t - 1) * 4] <= w);
  }
  lo = __reduce_min_sync(0xFFFFFFFFu, lo);
  hi = __reduce_max_sync(0xFFFFFFFFu, hi);
  sorted = __all_sync(0xFFFFFFFFu, sorted);
  if (!sorted) {
    lo = 0u;
    hi = 255u;
  }
  if (lane == 0) tile_q[k] = (uint16_t)(lo | (hi << 8));
}

// Pair tests against a tile whose first compare word is ascending: only the prefix with word_0 <= v[0] can dominate, so
// one binary search replaces that compare and the loop stops at the longest prefix of the warp (the 32 targets of a warp
// are neighbours in the same order, their prefixes are close).
template <int M, int W, int NV, int T, bool GID>
__device__ __forceinline__ int rank_pair_tests_sorted(const uint4* tb, const uint32_t* v, uint32_t gidv, int best) {
  int p = 0;  // number of records with word_0 <= v[0]
#pragma unroll
  for (int step = T / 2; step >= 1; step >>= 1)
    if (reinterpret_cast<const uint32_t*>(&tb[(p + step - 1) * NV])[0] <= v[0]) p += step;
  if (p < T && reinterpret_cast<const uint32_t*>(&tb[p * NV])[0] <= v[0]) ++p;
  const int pmax = __reduce_max_sync(0xFFFFFFFFu, p);
  int acc[4] = {best, 0, 0, 0};
#pragma unroll 8
  for (int s = 0; s < pmax; ++s) {
    uint32_t sw[W];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const uint4 a4 = tb[s * NV + q];
      sw[4 * q + 0] = a4.x;
      sw[4 * q + 1] = a4.y;
      sw[4 * q + 2] = a4.z;
      sw[4 * q + 3] = a4.w;
    }
    bool dom = (s < p) && (GID ? (sw[M - 1] != gidv) : true);
#pragma unroll
    for (int j = 1; j < M - 1; ++j) dom = dom && (sw[j] <= v[j]);
    acc[s & 3] = dom ? max(acc[s & 3], (int)sw[M]) : acc[s & 3];
  }
  return max(max(acc[0], acc[1]), max(acc[2], acc[3]));
}

// 128 pair tests of one streamed tile against this thread's record.  SKIP0: the first compare word is known to pass;
// GID: the "not identical" test is needed; C1: objective 1 must be compared too (source may share the target's segment)
template <int M, int W, int NV, int T, bool SKIP0, bool GID, bool C1>
__device__ __forceinline__ int rank_pair_tests(const uint4* tb, const uint32_t* c1tb, const uint32_t* v, uint32_t gidv,
                                               uint32_t c1v, int best) {
  int acc[4] = {best, 0, 0, 0};  // four independent max chains (the predicated max is the only loop-carried dependency)
#pragma unroll 16
  for (int s = 0; s < T; ++s) {
    uint32_t sw[W];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const uint4 a4 = tb[s * NV + q];
      sw[4 * q + 0] = a4.x;
      sw[4 * q + 1] = a4.y;
      sw[4 * q + 2] = a4.z;
      sw[4 * q + 3] = a4.w;
    }
    bool dom = GID ? (sw[M - 1] != gidv) : true;
#pragma unroll
    for (int j = SKIP0 ? 1 : 0; j < M - 1; ++j) dom = dom && (sw[j] <= v[j]);
    if (C1) dom = dom && (c1tb[s] <= c1v);
    const int r1 = (int)sw[M];
    acc[s & 3] = dom ? max(acc[s & 3], r1) : acc[s & 3];
  }
  return max(max(acc[0], acc[1]), max(acc[2], acc[3]));
}

// Breadth-first walk by path length from the node set `f` (level `level`): table[i * DLD + col] = last level at which
// i is reached = longest path.  Each thread walks its own source; the frontier is a 128-bit set in registers.
template <int T, int DLD>
__device__ __forceinline__ void frontier_fill(int8_t* table, const uint4* succ, uint4 f, int col, int level) {
  static_assert(T == 128, "frontier sets are 128 bits wide");
  while ((f.x | f.y | f.z | f.w) != 0u) {
    uint4 nx = make_uint4(0u, 0u, 0u, 0u);
#define DMO_WALK(word, base)                           \
  for (uint32_t mm = (word); mm != 0u; mm &= mm - 1u) { \
    const int i = (base) + __ffs(mm) - 1;               \
    table[i * DLD + col] = (int8_t)level;               \
    const uint4 sc = succ[i];                           \
    nx.x |= sc.x;                                       \
    nx.y |= sc.y;                                       \
    nx.z |= sc.z;                                       \
    nx.w |= sc.w;                                       \
  }
    DMO_WALK(f.x, 0)
    DMO_WALK(f.y, 32)
    DMO_WALK(f.z, 64)
    DMO_WALK(f.w, 96)
#undef DMO_WALK
    f = nx;
    ++level;
  }
}

// max(init, max_a (vals16[a] + row[a])) over one table row of T int8 path lengths (SENT = -128 = "no path"), two
// 16-bit lanes per instruction: PRMT + AND expand a byte pair to int16 (SENT becomes -32768, so SENT + value < 0 never
// wins against init >= 0), VIADDMNMX.S16x2 does the add and the max.  Requires 0 <= vals16, init <= 32127.
template <int T>
__device__ __forceinline__ int maxplus_packed(const int8_t* row, const int16_t* vals16, int init) {
  const uint4* dr = reinterpret_cast<const uint4*>(row);
  const uint4* vr = reinterpret_cast<const uint4*>(vals16);
  unsigned acc0 = ((unsigned)init & 0xFFFFu) * 0x00010001u, acc1 = acc0;
#define DMO_EXP_LO(w) (__byte_perm((w), 0u, 0x1100) & 0x807F807Fu)
#define DMO_EXP_HI(w) (__byte_perm((w), 0u, 0x3322) & 0x807F807Fu)
#pragma unroll
  for (int c = 0; c < T / 16; ++c) {
    const uint4 dv = dr[c];
    const uint4 v0 = vr[2 * c], v1 = vr[2 * c + 1];
    acc0 = __viaddmax_s16x2(DMO_EXP_LO(dv.x), v0.x, acc0);
    acc1 = __viaddmax_s16x2(DMO_EXP_HI(dv.x), v0.y, acc1);
    acc0 = __viaddmax_s16x2(DMO_EXP_LO(dv.y), v0.z, acc0);
    acc1 = __viaddmax_s16x2(DMO_EXP_HI(dv.y), v0.w, acc1);
    acc0 = __viaddmax_s16x2(DMO_EXP_LO(dv.z), v1.x, acc0);
    acc1 = __viaddmax_s16x2(DMO_EXP_HI(dv.z), v1.y, acc1);
    acc0 = __viaddmax_s16x2(DMO_EXP_LO(dv.w), v1.z, acc0);
    acc1 = __viaddmax_s16x2(DMO_EXP_HI(dv.w), v1.w, acc1);
  }
#undef DMO_EXP_LO
#undef DMO_EXP_HI
  const unsigned m = __vmaxs2(acc0, acc1);
  return max((int)(int16_t)(m & 0xFFFFu), (int)(int16_t)(m >> 16));
}

template <int M, int T, bool SEG>
__global__ void __launch_bounds__(T, SEG ? 4 : 5) rank_chain_kernel(uint32_t* rec, int nblocks, int* __restrict__ rankS, int* ticket,
                                                                 int* errflag, long long* trace, RankSeg sg) {
  // optional per-block time stamps (DMO_RANK_TRACE=<file>): 16 x globaltimer ns, then 16 x clock64, see scripts/rank_trace.py
#define RANK_TRACE(slot)                                                         \
  if (trace != nullptr && tid == 0) {                                            \
    long long gt_;                                                               \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                     \
    trace[(int64_t)b * 32 + (slot)] = gt_;                                       \
    trace[(int64_t)b * 32 + 16 + (slot)] = clock64();                             \
  }
  constexpr int W = 4 * ((M + 1 + 3) / 4);
  constexpr int NV = W / 4;
  constexpr int NW = T / 32;
  constexpr int RQ = M / 4, RC = M % 4;  // uint4 / component holding the rank word
  // Path-length tables are stored TRANSPOSED, [destination i][source], row stride T + 16 bytes: the thread that owns
  // destination i reads its whole row with 16-byte loads (stride 36 words -> the 8 threads of a quarter warp hit 8
  // distinct 4-word groups, conflict free), while the threads that fill the tables walk columns (consecutive bytes).
  constexpr int DLD = T + 16;
  constexpr int SENT = -128;  // "no in-block path"; real path lengths are 0 .. T-1 <= 127
  constexpr int PACK_LIMIT = 32000;  // ranks up to here take the packed 16-bit max-plus path
  __shared__ uint4 tile[T * NV];   // own block's records during table construction, then stream buffer 0
  constexpr bool DBUF = NV <= 2;  // M == 8 (three uint4 per record) would exceed 48 KB of static shared memory
  __shared__ uint4 tile2[DBUF ? T * NV : 1];  // stream buffer 1
  __shared__ __align__(16) int sh_r1[T];
  __shared__ __align__(16) int16_t sh_h16[T];
  __shared__ uint4 sh_succ[T];
  __shared__ __align__(16) int8_t sD[T * DLD];
  __shared__ __align__(16) int8_t sE[T * DLD];
  __shared__ int sh_blk;
  // segmented order only: objective-1 ids of the own block and of the streamed tiles, cached tile bands
  __shared__ uint32_t c1own[SEG ? T : 1];
  __shared__ uint32_t c1t[SEG ? 2 * T : 1];
  __shared__ uint16_t sq16[SEG ? RANK_SEG_MAXT : 1];
  __shared__ uint32_t sh_band[SEG ? 2 * NW : 1];
  __shared__ uint32_t skey[SEG ? T : 1];  // the block's last compare words in ascending order (staircase keys)

  const int tid = threadIdx.x;

  for (;;) {
    if (tid == 0) sh_blk = atomicAdd(ticket, 1);
    __syncthreads();
    const int b = sh_blk;
    __syncthreads();
    if (b >= nblocks) return;
    const int64_t i = (int64_t)b * T + tid;
    RANK_TRACE(0);
    if (trace != nullptr && tid == 0) {
      unsigned smid_;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));
      trace[(int64_t)b * 32 + 7] = smid_;
    }

    // ---- own record -> registers and shared tile
    uint32_t v[W];
    {
      const uint4* src = reinterpret_cast<const uint4*>(rec + i * W);
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        uint4 a = src[q];
        tile[tid * NV + q] = a;
        v[4 * q + 0] = a.x;
        v[4 * q + 1] = a.y;
        v[4 * q + 2] = a.z;
        v[4 * q + 3] = a.w;
      }
    }
    const uint32_t gidv = v[M - 1];
    uint32_t c1v = 0u;
    uint32_t wq_lo = 0u, wq_hi = 255u;  // band of this warp's 32 records (a quarter of the block's band)
    if (SEG) {
      c1v = sg.c1rec[i];
      c1own[tid] = c1v;
      const uint32_t q = min(v[0] >> sg.qshift, 255u);
      const uint32_t qlo = __reduce_min_sync(0xFFFFFFFFu, q), qhi = __reduce_max_sync(0xFFFFFFFFu, q);
      wq_lo = qlo;
      wq_hi = qhi;
      if ((tid & 31) == 0) {
        sh_band[tid >> 5] = qlo;
        sh_band[NW + (tid >> 5)] = qhi;
      }
      for (int t = tid; t < b - 1; t += T) sq16[t] = sg.tile_q[t];  // bands of every tile this block may stream
    }
    __syncthreads();
    const uint32_t first_gid = reinterpret_cast<const uint32_t*>(&tile[0])[M - 1];  // group of the block's first record
    uint32_t bq_lo = 0u, bq_hi = 255u;
    int kc = 0;  // tiles >= kc may hold sources of the block's own segment(s): full test incl. objective 1, never skipped
    int kg = 0;  // tiles >= kg may hold a copy of one of the block's vectors ("not identical" test needed)
    int spos = tid;  // position of this record in the block's staircase order
    if (SEG) {
      bq_lo = min(min(sh_band[0], sh_band[1]), min(sh_band[2], sh_band[3]));
      bq_hi = max(max(sh_band[NW], sh_band[NW + 1]), max(sh_band[NW + 2], sh_band[NW + 3]));
      kc = (int)(sg.seg_start[c1own[0] >> sg.sshift] / (uint32_t)T);
      {  // first position whose group id is >= the block's first group id (group ids grow along the order)
        int64_t lo = 0, hi = (int64_t)b * T;
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if (rec[mid * W + (M - 1)] < first_gid) lo = mid + 1; else hi = mid;
        }
        kg = (int)(lo / T);
      }
      if (M == 3) {  // rank of the own last compare word inside the block (ties by position): 128 compares per thread
        const uint32_t key = v[1];
        int cnt = 0;
#pragma unroll 8
        for (int s = 0; s < T; ++s) {
          const uint32_t ks = reinterpret_cast<const uint32_t*>(&tile[s * NV])[1];
          cnt += (ks < key || (ks == key && s < tid)) ? 1 : 0;
        }
        spos = cnt;
        skey[spos] = key;
      } else {
        skey[tid] = 0u;
      }
    }

    // ---- in-block successor bitmasks: succ[j] = { i > j in this block : j dominates i }; independent of any rank
    {
      uint32_t sm[NW];
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        uint32_t m = 0u;
#pragma unroll 8
        for (int s = 0; s < 32; ++s) {
          const uint32_t* sp = reinterpret_cast<const uint32_t*>(&tile[(w * 32 + s) * NV]);
          bool dom = (w * 32 + s > tid) && (sp[M - 1] != gidv);
#pragma unroll
          for (int j = 0; j < M - 1; ++j) dom = dom && (v[j] <= sp[j]);
          if (SEG) dom = dom && (c1v <= c1own[w * 32 + s]);  // objective 1 is not implied by the order inside a segment
          m |= (dom ? 1u : 0u) << s;
        }
        sm[w] = m;
      }
      static_assert(NW == 4, "successor masks are stored as one uint4 per node");
      sh_succ[tid] = make_uint4(sm[0], sm[1], sm[2], sm[3]);
      // the path-length table starts as "no path"
      const uint4 fill = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
      for (int t = tid; t < T * DLD / 16; t += T) reinterpret_cast<uint4*>(sD)[t] = fill;
    }
    __syncthreads();

    // ---- in-block longest-path table, computed BEFORE any rank is needed (off the critical path):
    // D[a][i] = number of edges of the longest in-block domination chain a -> ... -> i (SENT if none, 0 for i == a).
    // Thread a walks the DAG breadth first by path length: F_l = nodes reached from a by a path of exactly l edges,
    // F_{l+1} = union of succ[i] over i in F_l; the last level that contains i is the longest path (later writes win).
    // With it the in-block resolution is one max-plus product  rank_i = max_a (best_a + D[a][i]).
    frontier_fill<T, DLD>(sD, sh_succ, sh_succ[tid], tid, 1);
    sD[tid * DLD + tid] = 0;

    // ---- predecessor table: E[s][i] = longest in-block continuation of a chain that enters this block from point s of
    // block b-1 (0 if s dominates i directly).  Same walk, seeded with the block nodes s dominates, level 0.
    {
      const uint4 fill = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
      for (int t = tid; t < T * DLD / 16; t += T) reinterpret_cast<uint4*>(sE)[t] = fill;
    }
    __syncthreads();
    if (b > 0) {
      const int64_t ps = (int64_t)(b - 1) * T + tid;
      uint32_t pv[W];
      {
        const uint4* src = reinterpret_cast<const uint4*>(rec + ps * W);
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          uint4 a4 = src[q];  // static words only (ids, group); the rank word is read later with ld.relaxed
          pv[4 * q + 0] = a4.x;
          pv[4 * q + 1] = a4.y;
          pv[4 * q + 2] = a4.z;
          pv[4 * q + 3] = a4.w;
        }
      }
      const uint32_t pc1 = SEG ? sg.c1rec[ps] : 0u;
      uint32_t pm[NW];
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        uint32_t m = 0u;
#pragma unroll 8
        for (int s2 = 0; s2 < 32; ++s2) {
          const uint32_t* sp = reinterpret_cast<const uint32_t*>(&tile[(w * 32 + s2) * NV]);
          bool dom = (sp[M - 1] != pv[M - 1]);
#pragma unroll
          for (int j = 0; j < M - 1; ++j) dom = dom && (pv[j] <= sp[j]);
          if (SEG) dom = dom && (pc1 <= c1own[w * 32 + s2]);
          m |= (dom ? 1u : 0u) << s2;
        }
        pm[w] = m;
      }
      frontier_fill<T, DLD>(sE, sh_succ, make_uint4(pm[0], pm[1], pm[2], pm[3]), tid, 0);
    }
    __syncthreads();

    int best = 0;
    // ---- the last streamed tile (block b-2) is the second serial dependency: its ranks arrive one link before this
    // block's turn.  Its dominance pattern does not depend on ranks, so it is evaluated now into a 128-bit mask per
    // thread (kept in the successor-mask storage, which the table walks no longer need); when the ranks arrive only a
    // maximum over the set bits remains.
    bool sparse_b2 = false;
    if (SEG && b >= 2) {
      const int64_t p2 = (int64_t)(b - 2) * T + tid;
      tile2[tid] = *reinterpret_cast<const uint4*>(rec + p2 * W);  // static words; NV == 1 in the segmented kernels
      c1t[T + tid] = sg.c1rec[p2];
      __syncthreads();
      uint32_t mk[NW];
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        uint32_t m = 0u;
#pragma unroll 8
        for (int s2 = 0; s2 < 32; ++s2) {
          const uint32_t* sp = reinterpret_cast<const uint32_t*>(&tile2[w * 32 + s2]);
          bool dom = (sp[M - 1] != gidv) && (c1t[T + w * 32 + s2] <= c1v);
#pragma unroll
          for (int j = 0; j < M - 1; ++j) dom = dom && (sp[j] <= v[j]);
          m |= (dom ? 1u : 0u) << s2;
        }
        mk[w] = m;
      }
      sh_succ[tid] = make_uint4(mk[0], mk[1], mk[2], mk[3]);  // own slot only; all table walks finished at the barrier above
      // the set-bit walk only pays when the masks are sparse (a converged population: few dominators per point);
      // dense masks (random data) keep the pair tests, whose 128 steps pipeline better than a long dependent walk
      sparse_b2 = __syncthreads_or((__popc(mk[0]) + __popc(mk[1]) + __popc(mk[2]) + __popc(mk[3])) > 8 ? 1 : 0) == 0;
    }

    // ---- stream every earlier block except the predecessor: best = max over dominators of (rank + 1)
    RANK_TRACE(1);
    // Software pipelined: the data of the next tile (static words and, speculatively, its rank word -- or its staircase
    // entry) is requested before the current tile is evaluated, and the shared tile is double buffered, so a block that is
    // behind the wavefront pays one barrier and the evaluation per tile, not an L2 round trip on top; a block at the
    // wavefront only waits for the rank word.  In the segmented order whole tiles are skipped (see RankSeg).
    {
      auto next_tile = [&](int k) {
        if (SEG) {
          const int lim = min(kc, b - 1);
          while (k < lim && (uint32_t)(sq16[k] & 0xFFu) > bq_hi) ++k;  // every source word above the block's band
        }
        return k;
      };
      // tile kinds: 2 = staircase (earlier segment, every first word below the band, no copy of a block vector),
      //             1 = pair tests incl. objective 1 (may share a segment), 0 = pair tests,
      //             3 = the last tile: maximum over the precomputed dominance mask
      auto kind_of = [&](int k) {
        if (!SEG) return 0;
        if (k == b - 2 && sparse_b2) return 3;  // sparse dominance mask precomputed above
        if (k >= kc) return 1;
        return ((uint32_t)(sq16[k] >> 8) < bq_lo && k < kg) ? 2 : 0;
      };
      uint4 cur[NV];
      uint32_t cur_c1 = 0u;
      unsigned long long cur_st = 0ull;
      auto request = [&](int k, int kind) {
        if (SEG && kind == 2) {
          cur_st = ld_relaxed_u64(sg.stair + (int64_t)k * T + tid);
        } else {
          const uint4* src = reinterpret_cast<const uint4*>(rec + ((int64_t)k * T + tid) * W);
#pragma unroll
          for (int q = 0; q < NV; ++q) cur[q] = ld_relaxed_v4(src + q);
          if (SEG && kind == 1) cur_c1 = sg.c1rec[(int64_t)k * T + tid];
        }
      };
      int k = next_tile(0);
      int kind = k < b - 1 ? kind_of(k) : 0;
      if (k < b - 1) request(k, kind);
      int pb = 0;  // stream buffer parity
      while (k < b - 1) {
        if (k == b - 2) RANK_TRACE(2);
        {
          unsigned spins = 0;
          if (SEG && kind == 2) {
            const unsigned long long* src = sg.stair + (int64_t)k * T + tid;
            while ((cur_st >> 32) == 0ull) {  // the owner has not published its staircase yet
              __nanosleep(spins < 8 ? 100 : 400);
              cur_st = ld_relaxed_u64(src);
              if ((++spins & 0xFFu) == 0u && (spins > (1u << 21) || ld_relaxed_u32((const uint32_t*)errflag) != 0u)) {
                atomicExch(errflag, 1);
                break;
              }
            }
          } else {
            const uint4* src = reinterpret_cast<const uint4*>(rec + ((int64_t)k * T + tid) * W);
            while (word_of(cur[RQ], RC) == 0u) {  // not final yet: this block has caught up with the wavefront
              __nanosleep(spins < 8 ? 100 : 400);
              cur[RQ] = ld_relaxed_v4(src + RQ);
              if ((++spins & 0xFFu) == 0u && (spins > (1u << 21) || ld_relaxed_u32((const uint32_t*)errflag) != 0u)) {
                atomicExch(errflag, 1);
                break;
              }
            }
          }
        }
        uint4* tb = (DBUF && pb) ? tile2 : tile;
        if (!DBUF) __syncthreads();  // single buffer: everyone must be done with the previous tile
        if (SEG && kind == 2) {
          reinterpret_cast<unsigned long long*>(tb)[tid] = cur_st;
        } else {
#pragma unroll
          for (int q = 0; q < NV; ++q) tb[tid * NV + q] = cur[q];
          if (SEG && kind == 1) c1t[pb * T + tid] = cur_c1;
        }
        // Group ids grow along the order, so a tile whose last record is in an earlier group than this block's first
        // record holds no copy of any of this block's vectors: the "not identical" test can be dropped (one compare
        // per pair less).  The owner of the tile's last record votes through the tile barrier.
        const bool last_shares = !(SEG && kind == 2) && (tid == T - 1) && (word_of(cur[(M - 1) / 4], (M - 1) % 4) >= first_gid);
        const int kn = next_tile(k + 1);
        const int kind_n = kn < b - 1 ? kind_of(kn) : 0;
        if (kn < b - 1) request(kn, kind_n);  // consumed after this tile's evaluation
        // one barrier per tile: a stream buffer is rewritten two tiles later, after the next tile's barrier
        const bool may_share_group = __syncthreads_or(last_shares ? 1 : 0) != 0;
        if (k == b - 2) RANK_TRACE(3);
        const uint32_t* c1tb = c1t + (SEG ? pb * T : 0);
        if (SEG && kind == 2) {
          const uint2* st = reinterpret_cast<const uint2*>(tb);  // .x = key (ascending), .y = running max of rank + 1
          if (M == 3) {
            int idx = 0;  // number of keys <= v[1]
#pragma unroll
            for (int step = T / 2; step >= 1; step >>= 1)
              if (st[idx + step - 1].x <= v[1]) idx += step;
            if (idx < T && st[idx].x <= v[1]) ++idx;  // T is a power of two: the steps cover T - 1 positions
            if (idx > 0) best = max(best, (int)st[idx - 1].y);
          } else {
            best = max(best, (int)st[T - 1].y);  // two objectives: every record of the tile dominates the block
          }
        } else if (SEG && kind == 3) {
          const uint4 mk4 = sh_succ[tid];
          const uint32_t mk[4] = {mk4.x, mk4.y, mk4.z, mk4.w};
#pragma unroll
          for (int w = 0; w < 4; ++w)
            for (uint32_t mm = mk[w]; mm != 0u; mm &= mm - 1u) {
              const int s2 = w * 32 + __ffs(mm) - 1;
              best = max(best, (int)reinterpret_cast<const uint32_t*>(&tb[s2 * NV])[M]);
            }
        } else if (kind == 1) {
          best = rank_pair_tests<M, W, NV, T, false, true, true>(tb, c1tb, v, gidv, c1v, best);
        } else if (SEG && (uint32_t)(sq16[k] & 0xFFu) > wq_hi) {
          // straddles the block's band but lies entirely above this warp's: nothing in it dominates these 32 records
        } else if (SEG && (uint32_t)(sq16[k] >> 8) < wq_lo) {  // every source word below this warp's band
          best = may_share_group ? rank_pair_tests<M, W, NV, T, true, true, false>(tb, c1tb, v, gidv, c1v, best)
                                 : rank_pair_tests<M, W, NV, T, true, false, false>(tb, c1tb, v, gidv, c1v, best);
        } else if (SEG && sq16[k] != 0xFF00u) {  // ascending first word (the full band marks the tiles that are not)
          best = may_share_group ? rank_pair_tests_sorted<M, W, NV, T, true>(tb, v, gidv, best)
                                 : rank_pair_tests_sorted<M, W, NV, T, false>(tb, v, gidv, best);
        } else {
          best = may_share_group ? rank_pair_tests<M, W, NV, T, false, true, false>(tb, c1tb, v, gidv, c1v, best)
                                 : rank_pair_tests<M, W, NV, T, false, false, false>(tb, c1tb, v, gidv, c1v, best);
        }
        // Second barrier of the tile.  The double buffering alone orders the accesses (a buffer is rewritten two tiles
        // later, after the next tile's barrier, which every reader of this tile has to reach first), but compute-sanitizer's
        // racecheck reports the read above against the next write of the same buffer as a potential WAR hazard; with this
        // barrier the tool is clean (profiles/r2_sanitizer.txt) at no measurable cost (the chain is latency bound).
        __syncthreads();
        k = kn;
        kind = kind_n;
        pb ^= 1;
      }
    }
    // ---- in-block resolution, part 1 (before the predecessor's ranks are needed):
    //   Rb_i = max_a (bulk_a + D[a][i]) folds the contributions of all blocks < b-1 through the in-block paths.
    int r = best;
    if (!__syncthreads_or(best > PACK_LIMIT)) {
      sh_h16[tid] = (int16_t)best;
      __syncthreads();
      r = maxplus_packed<T>(sD + tid * DLD, sh_h16, best);
    } else {  // more than 32000 fronts: plain 32-bit max-plus over the same table
      sh_r1[tid] = best;
      __syncthreads();
      for (int a2 = 0; a2 < T; ++a2) {
        const int dl = (int)sD[tid * DLD + a2];
        r = (dl >= 0) ? max(r, sh_r1[a2] + dl) : r;
      }
    }
    __syncthreads();

    // ---- the predecessor block is the critical dependency.  Everything that does not depend on its ranks is done
    // first: its dominance pattern (pmask) and, from it, the table  E[s][i] = longest in-block continuation of a chain that
    // enters this block from predecessor point s and ends at i  (E = max_{a : s dominates a} D[a][i], SENT if none).
    // Once the ranks r1_s = rank_s + 1 arrive the resolution is a single max-plus product
    //     rank_i = max(Rb_i, max_s (r1_s + E[s][i])),
    // i.e. the critical path per block is: poll -> one barrier -> 128 independent loads / adds / maxes -> publish.
    if (b > 0) {
      const int k = b - 1;
      // ---- critical section starts here
      RANK_TRACE(4);
      bool big;
      {
        const uint32_t* rw = rec + ((int64_t)k * T + tid) * W + M;
        uint32_t r1 = ld_relaxed_u32(rw);
        unsigned spins = 0;
        while (r1 == 0u) {
          r1 = ld_relaxed_u32(rw);
          if ((++spins & 0xFFFFu) == 0u && (spins > (1u << 24) || ld_relaxed_u32((const uint32_t*)errflag) != 0u)) {
            atomicExch(errflag, 1);
            break;
          }
        }
        sh_r1[tid] = (int)r1;
        sh_h16[tid] = (int16_t)r1;
        big = (int)r1 > PACK_LIMIT || r > PACK_LIMIT;
      }
      const int any_big = __syncthreads_or(big ? 1 : 0);
      RANK_TRACE(5);
      if (!any_big) {
        r = maxplus_packed<T>(sE + tid * DLD, sh_h16, r);
      } else {
        for (int s2 = 0; s2 < T; ++s2) {
          const int e = (int)sE[tid * DLD + s2];
          r = (e >= 0) ? max(r, sh_r1[s2] + e) : r;
        }
      }
    }
    st_relaxed_u32(rec + i * W + M, (uint32_t)(r + 1));  // publish: the rank word doubles as the ready flag
    rankS[i] = r;
    RANK_TRACE(6);
    if (SEG) {
      // staircase of this tile for the blocks of later segments (off the critical path: they are at least a segment
      // away): running maximum of rank + 1 in ascending order of the last compare word; key and value travel in one
      // 64-bit store, a non-zero value marks the entry as published
      __syncthreads();
      sh_r1[spos] = r + 1;
      __syncthreads();
      int val = sh_r1[tid];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, val, off);
        if ((tid & 31) >= off) val = max(val, o);
      }
      if ((tid & 31) == 31) sh_band[tid >> 5] = (uint32_t)val;
      __syncthreads();
      for (int w = 0; w < (tid >> 5); ++w) val = max(val, (int)sh_band[w]);
      st_relaxed_u64(sg.stair + i, ((unsigned long long)(uint32_t)val << 32) | (unsigned long long)skey[tid]);
    }
    __syncthreads();
  }
#undef RANK_TRACE
}

// Rank-0 test only (filter for the hypervolume / EHVI routines): "is some point dominating me" has no dependency chain,
// so it is the bulk scan alone, with a block-wide early exit once every target of the block has found a dominator.
template <int M, int T>
__global__ void __launch_bounds__(T) nd_flag_kernel(const uint32_t* __restrict__ rec, int nblocks, int* __restrict__ flagS) {
  constexpr int W = 4 * ((M + 1 + 3) / 4);
  constexpr int NV = W / 4;
  __shared__ uint4 tile[T * NV];
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int64_t i = (int64_t)b * T + tid;
  uint32_t v[W];
  {
    const uint4* src = reinterpret_cast<const uint4*>(rec + i * W);
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      uint4 a = src[q];
      v[4 * q + 0] = a.x;
      v[4 * q + 1] = a.y;
      v[4 * q + 2] = a.z;
      v[4 * q + 3] = a.w;
    }
  }
  const uint32_t gidv = v[M - 1];
  bool dominated = false;
  for (int k = 0; k <= b; ++k) {
    if (__syncthreads_and(dominated ? 1 : 0)) break;
    {
      const uint4* src = reinterpret_cast<const uint4*>(rec + (int64_t)k * T * W);
#pragma unroll
      for (int q = 0; q < NV; ++q) tile[tid * NV + q] = src[tid * NV + q];
    }
    __syncthreads();
    const int lim = (k == b) ? tid : T;  // inside the own block only earlier positions can dominate
#pragma unroll 8
    for (int s = 0; s < T; ++s) {
      uint32_t sw[W];
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        uint4 a = tile[s * NV + q];
        sw[4 * q + 0] = a.x;
        sw[4 * q + 1] = a.y;
        sw[4 * q + 2] = a.z;
        sw[4 * q + 3] = a.w;
      }
      bool dom = (s < lim) && (sw[M - 1] != gidv);
#pragma unroll
      for (int j = 0; j < M - 1; ++j) dom = dom && (sw[j] <= v[j]);
      dominated = dominated || dom;
    }
  }
  flagS[i] = dominated ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ rank-0 test on a cell grid
// For two and three objectives the records carry one or two compare words (the first objective is implied by the
// lexicographic order).  A G x G grid over those words answers most rank-0 queries with one table lookup: a point in a
// cell strictly below the target's cell in both words, and earlier in lexicographic order, dominates it, so the target
// is dominated iff the minimum position over those cells (an exclusive 2-D prefix minimum of the per-cell minima) is
// smaller than its own.  Only the points in the target's own cell row and cell column need the exact test; two
// cell-ordered copies of the records (row-major and column-major) make both of them contiguous streams.
// (M == 2 duplicates its single compare word, i.e. only the diagonal cells are populated.)
// cell of a dense id: the ids of an objective run from 0 to maxid (its number of distinct values - 1), which can be far below
// n (quantised or heavily tied objectives); the shift follows maxid, so the grid stays populated evenly either way
__device__ __forceinline__ int cell_shift(uint32_t maxid, int gbits) {
  const int b = 32 - __clz(maxid | 1u);
  return b > gbits ? b - gbits : 0;
}

__global__ void ndg_key_kernel(const uint32_t* __restrict__ rec, int64_t npad, int M, const uint32_t* __restrict__ maxid, int gbits,
                               uint32_t* __restrict__ keyA, uint32_t* __restrict__ keyB, uint32_t* __restrict__ pos) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npad) return;
  const uint32_t G1 = (1u << gbits) - 1u;
  const uint32_t* w = rec + p * 4;  // W == 4 for M <= 3
  const uint32_t a = min(w[0] >> cell_shift(maxid[1], gbits), G1);
  const uint32_t b = (M == 3) ? min(w[1] >> cell_shift(maxid[2], gbits), G1) : a;
  keyA[p] = (a << gbits) | b;
  keyB[p] = (b << gbits) | a;
  pos[p] = (uint32_t)p;
}

// cell-ordered copy: (word 0, word 1 (word 0 again for M == 2), group id, lexicographic position)
__global__ void ndg_gather_kernel(const uint32_t* __restrict__ rec, const uint32_t* __restrict__ order, int64_t npad,
                                  int M, uint4* __restrict__ crec) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npad) return;
  const uint32_t p = order[t];
  const uint4 r = *reinterpret_cast<const uint4*>(rec + (int64_t)p * 4);
  crec[t] = (M == 3) ? make_uint4(r.x, r.y, r.z, p) : make_uint4(r.x, r.x, r.y, p);
}

// cstart[c] = first slot whose key is >= c, c = 0 .. ncell (lower bounds over the sorted keys)
__global__ void ndg_start_kernel(const uint32_t* __restrict__ skey, int64_t npad, int ncell, uint32_t* __restrict__ cstart) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > ncell) return;
  int64_t lo = 0, hi = npad;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (skey[mid] < (uint32_t)c) lo = mid + 1; else hi = mid;
  }
  cstart[c] = (uint32_t)lo;
}

// inclusive prefix minimum of the per-cell minimum positions: pass 0 along each row, pass 1 along each column
__global__ void ndg_prefix_min_kernel(const uint32_t* __restrict__ cstartA, const uint4* __restrict__ crecA, int gbits,
                                      int pass, uint32_t* __restrict__ pm) {
  extern __shared__ uint32_t sh_scan[];
  const int G = 1 << gbits;
  const int line = blockIdx.x, t = threadIdx.x;  // pass 0: line = row a, t = column b; pass 1: line = column b, t = row a
  const int cell = pass == 0 ? line * G + t : t * G + line;
  uint32_t v;
  if (pass == 0) {
    const uint32_t s0 = cstartA[cell], s1 = cstartA[cell + 1];
    v = s0 < s1 ? crecA[s0].w : 0xFFFFFFFFu;  // stable sort: the first record of a cell has its smallest position
  } else {
    v = pm[cell];
  }
  sh_scan[t] = v;
  __syncthreads();
  for (int off = 1; off < G; off <<= 1) {
    const uint32_t o = t >= off ? sh_scan[t - off] : 0xFFFFFFFFu;
    __syncthreads();
    v = min(v, o);
    sh_scan[t] = v;
    __syncthreads();
  }
  pm[cell] = v;
}

__global__ void ndg_flag_kernel(const uint32_t* __restrict__ rec, int64_t n, int M, const uint32_t* __restrict__ maxid, int gbits,
                                const uint32_t* __restrict__ pm, const uint32_t* __restrict__ cstartA,
                                const uint32_t* __restrict__ cstartB, const uint4* __restrict__ crecA,
                                const uint4* __restrict__ crecB, int* __restrict__ flagS) {
  const int64_t p64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p64 >= n) return;
  const uint32_t p = (uint32_t)p64;
  const int G = 1 << gbits;
  const uint4 r = *reinterpret_cast<const uint4*>(rec + p64 * 4);
  const uint32_t c0 = r.x, c1 = (M == 3) ? r.y : r.x, gid = (M == 3) ? r.z : r.y;
  const int sh0 = cell_shift(maxid[1], gbits), sh1 = (M == 3) ? cell_shift(maxid[2], gbits) : sh0;
  const int a = (int)min(c0 >> sh0, (uint32_t)(G - 1)), b = (int)min(c1 >> sh1, (uint32_t)(G - 1));
  bool dom = a > 0 && b > 0 && __ldg(pm + (a - 1) * G + (b - 1)) < p;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if (dom) break;
    const uint4* cr = pass == 0 ? crecA : crecB;
    uint32_t t = pass == 0 ? __ldg(cstartA + a * G) : __ldg(cstartB + b * G);
    const uint32_t t1 = pass == 0 ? __ldg(cstartA + a * G + b + 1) : __ldg(cstartB + b * G + a);
#define DMO_NDG_TEST(q) ((q).w < p && (q).x <= c0 && (q).y <= c1 && (q).z != gid)
    for (; t + 4 <= t1 && !dom; t += 4) {
      const uint4 q0 = __ldg(cr + t), q1 = __ldg(cr + t + 1), q2 = __ldg(cr + t + 2), q3 = __ldg(cr + t + 3);
      dom = DMO_NDG_TEST(q0) || DMO_NDG_TEST(q1) || DMO_NDG_TEST(q2) || DMO_NDG_TEST(q3);
    }
    for (; t < t1 && !dom; ++t) {
      const uint4 q0 = __ldg(cr + t);
      dom = DMO_NDG_TEST(q0);
    }
#undef DMO_NDG_TEST
  }
  flagS[p64] = dom ? 1 : 0;
}

int bits_for(int64_t n);

// flagS[p] = 1 iff the record at lexicographic position p is dominated (M <= 3, W == 4)
int nd_flags_grid(dmo_ctx* ctx, const uint32_t* rec, int64_t n, int64_t npad, int M, const uint32_t* maxid, int* flagS) {
  const int bits = bits_for(n);
  int gbits = bits / 2;
  if (gbits < 4) gbits = 4;
  if (gbits > 9) gbits = 9;
  const int G = 1 << gbits, GG = G * G;
  DevBuf<uint32_t> keyA, keyB, keyS, pos, ord, cstartA, cstartB, pm;
  DevBuf<uint4> crecA, crecB;
  DMO_TRY(keyA.alloc(ctx, npad));
  DMO_TRY(keyB.alloc(ctx, npad));
  DMO_TRY(keyS.alloc(ctx, npad));
  DMO_TRY(pos.alloc(ctx, npad));
  DMO_TRY(ord.alloc(ctx, npad));
  DMO_TRY(cstartA.alloc(ctx, GG + 1));
  DMO_TRY(cstartB.alloc(ctx, GG + 1));
  DMO_TRY(pm.alloc(ctx, GG));
  DMO_TRY(crecA.alloc(ctx, npad));
  DMO_TRY(crecB.alloc(ctx, npad));
  const unsigned gp = (unsigned)ceil_div(npad, 256);
  ProfileScope ps(ctx, "nd_flags");
  DMO_LAUNCH(ndg_key_kernel, gp, 256, 0, rec, npad, M, maxid, gbits, keyA.p, keyB.p, pos.p);
  for (int pass = 0; pass < 2; ++pass) {  // stable sorts: positions stay ascending inside a cell
    DMO_TRY(prim_sort_pairs_u32(ctx, pass == 0 ? keyA.p : keyB.p, keyS.p, pos.p, ord.p, npad, 0, 2 * gbits));
    DMO_LAUNCH(ndg_gather_kernel, gp, 256, 0, rec, ord.p, npad, M, pass == 0 ? crecA.p : crecB.p);
    DMO_LAUNCH(ndg_start_kernel, (unsigned)ceil_div(GG + 1, 256), 256, 0, keyS.p, npad, GG,
               pass == 0 ? cstartA.p : cstartB.p);
  }
  DMO_LAUNCH(ndg_prefix_min_kernel, G, G, G * sizeof(uint32_t), cstartA.p, crecA.p, gbits, 0, pm.p);
  DMO_LAUNCH(ndg_prefix_min_kernel, G, G, G * sizeof(uint32_t), cstartA.p, crecA.p, gbits, 1, pm.p);
  DMO_LAUNCH(ndg_flag_kernel, (unsigned)ceil_div(n, 128), 128, 0, rec, n, M, maxid, gbits, pm.p, cstartA.p, cstartB.p,
             crecA.p, crecB.p, flagS);
  DMO_CHECK_LAUNCH();
  return DMO_OK;
}

template <int M>
int launch_nd_flags(dmo_ctx* ctx, const uint32_t* rec, int nblocks, int* flagS) {
  ProfileScope ps(ctx, "nd_flags");
  DMO_LAUNCH((nd_flag_kernel<M, RANK_T>), nblocks, RANK_T, 0, rec, nblocks, flagS);
  DMO_CHECK_LAUNCH();
  return DMO_OK;
}

__global__ void scatter_rank_kernel(const int* __restrict__ rankS, const uint32_t* __restrict__ perm, int64_t n,
                                    int32_t* __restrict__ rank) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) rank[perm[p]] = rankS[p];
}

__global__ void copy_u32_to_i32_kernel(const uint32_t* __restrict__ a, int64_t n, int32_t* __restrict__ out) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) out[p] = (int32_t)a[p];
}

template <int M, bool SEG>
int launch_chain(dmo_ctx* ctx, uint32_t* rec, int nblocks, int* rankS, int* ticket, int* errflag, const RankSeg& sg) {
  // debugging aid: DMO_RANK_TRACE=<file> dumps 32 int64 time stamps per block of the chain kernel
  DevBuf<long long> trace;
  const char* trace_path = getenv("DMO_RANK_TRACE");
  if (trace_path && *trace_path) {
    DMO_TRY(trace.alloc(ctx, (size_t)nblocks * 32));
    DMO_CUDA(cudaMemsetAsync(trace.p, 0, (size_t)nblocks * 32 * sizeof(long long), ctx->stream));
  }
  int occ = 0;
  DMO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rank_chain_kernel<M, RANK_T, SEG>, RANK_T, 0));
  if (occ < 1) occ = 1;
  // fewer co-resident CTAs per SM shorten the serial chain (the block on the critical path shares its SM's issue
  // slots with the others); DMO_RANK_OCC overrides for tuning
  int cap = 6;
  if (const char* e = getenv("DMO_RANK_OCC")) cap = atoi(e);
  if (cap >= 1 && occ > cap) occ = cap;
  int nctas = nblocks < occ * ctx->sm_count ? nblocks : occ * ctx->sm_count;
  {
    ProfileScope ps(ctx, "rank_chain");
    DMO_LAUNCH((rank_chain_kernel<M, RANK_T, SEG>), nctas, RANK_T, 0, rec, nblocks, rankS, ticket, errflag, trace.p, sg);
    DMO_CHECK_LAUNCH();
  }
  if (trace.p) {
    std::vector<long long> h((size_t)nblocks * 32);
    DMO_CUDA(cudaMemcpyAsync(h.data(), trace.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    DMO_CUDA(cudaStreamSynchronize(ctx->stream));
    if (FILE* f = fopen(trace_path, "wb")) {
      fwrite(h.data(), sizeof(long long), h.size(), f);
      fclose(f);
    }
  }
  return DMO_OK;
}

int bits_for(int64_t n) {
  int b = 1;
  while (((int64_t)1 << b) < n) ++b;
  return b;
}

// ------------------------------------------------------------------------------------------------ front peeling (M == 3)
// A truncation (remove_worst: keep the best `keep` of n rows) needs the ranks of the kept rows only.  When those rows span
// few fronts -- a converging population: the bench's merged sets put the best 65 536 of 131 072 points into 7 fronts, a
// sphere-shaped set into 2 -- peeling them one by one is cheaper than the chain, whose 1024 links are serial whatever the
// data looks like.  Front k = the points of the remaining set that no remaining point dominates, found with the cell grid
// of the rank-0 filter above, built once on the dense ids (cells over objectives 2 and 3; inside a cell the records are
// ordered by their objective-1 id, so the smallest living id of a cell is its first living record):
//   * per peel: per-cell minimum of the living objective-1 ids, its 2-D prefix minimum, one pass over the living points
//     (table lookup for the cells strictly below, exact tests along the own cell row and column), then the new front is
//     marked: rank written, its records in both cell-ordered copies overwritten with an id that dominates nothing;
//   * the loop stops once `keep` rows are ranked (the others get the next rank: they are truncated away), or gives up
//     when the fronts turn out to be small (many peels ahead): the chain then runs as if nothing had happened.
constexpr uint32_t PEEL_DEAD = 0xFFFFFFFFu;

__global__ void peel_key_kernel(const uint32_t* __restrict__ R, int64_t n, const uint32_t* __restrict__ maxid, int gbits,
                                uint32_t* __restrict__ key0, uint32_t* __restrict__ keyA, uint32_t* __restrict__ keyB) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t G1 = (1u << gbits) - 1u;
  const uint32_t a = min(R[n + i] >> cell_shift(maxid[1], gbits), G1), b = min(R[2 * n + i] >> cell_shift(maxid[2], gbits), G1);
  key0[i] = R[i];
  keyA[i] = (a << gbits) | b;
  keyB[i] = (b << gbits) | a;
}

// cell-ordered copy (ids of the three objectives, point index) and the slot of every point in it
__global__ void peel_gather_kernel(const uint32_t* __restrict__ R, const uint32_t* __restrict__ order, int64_t n,
                                   uint4* __restrict__ crec, uint32_t* __restrict__ slot) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t i = order[t];
  crec[t] = make_uint4(R[i], R[n + i], R[2 * n + i], i);
  slot[i] = (uint32_t)t;
}

// smallest living objective-1 id of every cell (the pointer to a cell's first living record only ever moves forward)
__global__ void peel_cellmin_kernel(const uint32_t* __restrict__ cstart, const uint4* __restrict__ crec, int ncell,
                                    uint32_t* __restrict__ first, uint32_t* __restrict__ pm) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  uint32_t f = first[c];
  const uint32_t e = cstart[c + 1];
  while (f < e && crec[f].x == PEEL_DEAD) ++f;
  first[c] = f;
  pm[c] = f < e ? crec[f].x : PEEL_DEAD;
}

// in-place inclusive prefix minimum: pass 0 along each row of the G x G table, pass 1 along each column
__global__ void peel_prefix_min_kernel(uint32_t* __restrict__ pm, int gbits, int pass) {
  extern __shared__ uint32_t sh_scan[];
  const int G = 1 << gbits;
  const int line = blockIdx.x, t = threadIdx.x;
  const int cell = pass == 0 ? line * G + t : t * G + line;
  uint32_t v = pm[cell];
  sh_scan[t] = v;
  __syncthreads();
  for (int off = 1; off < G; off <<= 1) {
    const uint32_t o = t >= off ? sh_scan[t - off] : PEEL_DEAD;
    __syncthreads();
    v = min(v, o);
    sh_scan[t] = v;
    __syncthreads();
  }
  pm[cell] = v;
}

// dom[i] = 1 iff a living point dominates the living point i
__global__ void peel_flag_kernel(const uint32_t* __restrict__ R, int64_t n, const uint32_t* __restrict__ maxid, int gbits,
                                 const uint32_t* __restrict__ pm,
                                 const uint32_t* __restrict__ cstartA, const uint32_t* __restrict__ cstartB,
                                 const uint4* __restrict__ crecA, const uint4* __restrict__ crecB,
                                 const uint8_t* __restrict__ alive, uint8_t* __restrict__ dom_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !alive[i]) return;
  const int G = 1 << gbits;
  const uint32_t c0 = R[i], c1 = R[n + i], c2 = R[2 * n + i];
  const int a = (int)min(c1 >> cell_shift(maxid[1], gbits), (uint32_t)(G - 1));
  const int b = (int)min(c2 >> cell_shift(maxid[2], gbits), (uint32_t)(G - 1));
  // cells strictly below in both words hold different vectors: "<=" on the first objective is enough there
  bool dom = a > 0 && b > 0 && __ldg(pm + (a - 1) * G + (b - 1)) <= c0;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if (dom) break;
    const uint4* cr = pass == 0 ? crecA : crecB;
    uint32_t t = pass == 0 ? __ldg(cstartA + a * G) : __ldg(cstartB + b * G);
    const uint32_t t1 = pass == 0 ? __ldg(cstartA + a * G + b + 1) : __ldg(cstartB + b * G + a);
    // a dead record carries PEEL_DEAD as its first id and fails the first compare
#define DMO_PEEL_TEST(q) ((q).x <= c0 && (q).y <= c1 && (q).z <= c2 && !((q).x == c0 && (q).y == c1 && (q).z == c2))
    // the few threads that get here walk hundreds of records: sixteen loads in flight per step (ncu: 9 % of the issue slots
    // busy with four, the kernel's time is the dependent-load latency of its longest walks)
    for (; t + 16 <= t1 && !dom; t += 16) {
      uint4 q[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) q[u] = __ldg(cr + t + u);
#pragma unroll
      for (int u = 0; u < 16; ++u) dom = dom || DMO_PEEL_TEST(q[u]);
    }
    for (; t + 4 <= t1 && !dom; t += 4) {
      const uint4 q0 = __ldg(cr + t), q1 = __ldg(cr + t + 1), q2 = __ldg(cr + t + 2), q3 = __ldg(cr + t + 3);
      dom = DMO_PEEL_TEST(q0) || DMO_PEEL_TEST(q1) || DMO_PEEL_TEST(q2) || DMO_PEEL_TEST(q3);
    }
    for (; t < t1 && !dom; ++t) {
      const uint4 q0 = __ldg(cr + t);
      dom = DMO_PEEL_TEST(q0);
    }
#undef DMO_PEEL_TEST
  }
  dom_out[i] = dom ? 1 : 0;
}

// the living points nobody dominates form front k: rank, death, count
__global__ void peel_mark_kernel(int64_t n, uint8_t* __restrict__ alive, const uint8_t* __restrict__ dom, int k,
                                 int32_t* __restrict__ rank, const uint32_t* __restrict__ slotA, const uint32_t* __restrict__ slotB,
                                 uint4* __restrict__ crecA, uint4* __restrict__ crecB, unsigned long long* __restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool mine = false;
  if (i < n && alive[i] && !dom[i]) {
    mine = true;
    rank[i] = k;
    alive[i] = 0;
    crecA[slotA[i]].x = PEEL_DEAD;
    crecB[slotB[i]].x = PEEL_DEAD;
  }
  const unsigned m = __ballot_sync(0xFFFFFFFFu, mine);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}

__global__ void peel_rest_kernel(int64_t n, const uint8_t* __restrict__ alive, int k, int32_t* __restrict__ rank) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && alive[i]) rank[i] = k;
}

// Is the first front worth peeling?  PEEL_PROBE evenly spaced points are tested against the whole set (every thread brings
// one point and runs it past the probes in shared memory); a probe that nobody dominates stands for n / PEEL_PROBE points of
// front 0.  A uniform cloud (front 0 = 0.05 % of the set) almost never passes, a converging population always does.
constexpr int PEEL_PROBE = 256;
__global__ void peel_probe_kernel(const uint32_t* __restrict__ R, int64_t n, unsigned* __restrict__ dominated) {
  __shared__ uint32_t sp[PEEL_PROBE][3];
  __shared__ unsigned sdom[PEEL_PROBE / 32];
  const int64_t stride = n / PEEL_PROBE;
  for (int t = threadIdx.x; t < PEEL_PROBE; t += blockDim.x) {
    const int64_t i = (int64_t)t * stride + (stride >> 1);
    sp[t][0] = R[i];
    sp[t][1] = R[n + i];
    sp[t][2] = R[2 * n + i];
  }
  if (threadIdx.x < PEEL_PROBE / 32) sdom[threadIdx.x] = 0u;
  __syncthreads();
  unsigned mine[PEEL_PROBE / 32];
#pragma unroll
  for (int w = 0; w < PEEL_PROBE / 32; ++w) mine[w] = 0u;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t a = R[j], b = R[n + j], c = R[2 * n + j];
#pragma unroll
    for (int w = 0; w < PEEL_PROBE / 32; ++w) {
      unsigned m = 0u;
#pragma unroll 8
      for (int t = 0; t < 32; ++t) {
        const uint32_t pa = sp[w * 32 + t][0], pb = sp[w * 32 + t][1], pc = sp[w * 32 + t][2];
        const bool d = a <= pa && b <= pb && c <= pc && !(a == pa && b == pb && c == pc);
        m |= d ? (1u << t) : 0u;
      }
      mine[w] |= m;
    }
  }
#pragma unroll
  for (int w = 0; w < PEEL_PROBE / 32; ++w) {
    const unsigned m = __reduce_or_sync(0xFFFFFFFFu, mine[w]);
    if ((threadIdx.x & 31) == 0 && m) atomicOr(&sdom[w], m);
  }
  __syncthreads();
  if (threadIdx.x < PEEL_PROBE / 32 && sdom[threadIdx.x]) atomicOr(&dominated[threadIdx.x], sdom[threadIdx.x]);
}

__global__ void fill_u8_kernel(uint8_t* __restrict__ a, int64_t n, uint8_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

// *done = true: d_rank holds exact ranks for (at least) the best `keep` rows, a common larger rank for the rest
int rank_by_peeling(dmo_ctx* ctx, const uint32_t* R, const uint32_t* maxid, int64_t n, int64_t keep, int32_t* d_rank, bool* done) {
  *done = false;
  int max_peels = 14;  // the chain costs about as much as 20 peels plus the grid
  if (const char* e = getenv("DMO_RANK_PEEL")) max_peels = atoi(e);
  if (max_peels <= 0) return DMO_OK;
  ProfileScope ps(ctx, "rank_peel");
  {  // probe before building anything: fewer than two undominated probes = a first front below ~1 % of the set
    DevBuf<unsigned> pd;
    DMO_TRY(pd.alloc(ctx, PEEL_PROBE / 32));
    DMO_CUDA(cudaMemsetAsync(pd.p, 0, (PEEL_PROBE / 32) * sizeof(unsigned), ctx->stream));
    const int gridp = (int)(ceil_div(n, 256) < 2 * (int64_t)ctx->sm_count ? ceil_div(n, 256) : 2 * (int64_t)ctx->sm_count);
    DMO_LAUNCH(peel_probe_kernel, gridp, 256, 0, R, n, pd.p);
    DMO_CHECK_LAUNCH();
    unsigned h[PEEL_PROBE / 32];
    DMO_CUDA(cudaMemcpyAsync(h, pd.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    DMO_CUDA(cudaStreamSynchronize(ctx->stream));
    int free_probes = PEEL_PROBE;
    for (int w = 0; w < PEEL_PROBE / 32; ++w) free_probes -= __builtin_popcount(h[w]);
    if (free_probes < 2 && !(getenv("DMO_RANK_PEEL_NOPROBE") && atoi(getenv("DMO_RANK_PEEL_NOPROBE")))) return DMO_OK;
  }
  const int bits = bits_for(n);
  int gbits = (bits + 1) / 2;  // two cells per point at n = 131 072: measured 0.80 ms per bench step against 0.96 ms with 2^8 cells per axis
  if (gbits < 4) gbits = 4;
  if (const char* e = getenv("DMO_PEEL_GBITS")) gbits = atoi(e);
  if (gbits > 9) gbits = 9;
  const int G = 1 << gbits, GG = G * G;
  const unsigned g = (unsigned)ceil_div(n, 256);
  DevBuf<uint32_t> key0, keyA, keyB, keyS, keyT, ord0, ordS, iota, cstartA, cstartB, firstA, pm, slotA, slotB;
  DevBuf<uint4> crecA, crecB;
  DevBuf<uint8_t> alive, dom;
  DevBuf<unsigned long long> count;
  DMO_TRY(key0.alloc(ctx, n));
  DMO_TRY(keyA.alloc(ctx, n));
  DMO_TRY(keyB.alloc(ctx, n));
  DMO_TRY(keyS.alloc(ctx, n));
  DMO_TRY(keyT.alloc(ctx, n));
  DMO_TRY(ord0.alloc(ctx, n));
  DMO_TRY(ordS.alloc(ctx, n));
  DMO_TRY(iota.alloc(ctx, n));
  DMO_TRY(cstartA.alloc(ctx, GG + 1));
  DMO_TRY(cstartB.alloc(ctx, GG + 1));
  DMO_TRY(firstA.alloc(ctx, GG));
  DMO_TRY(pm.alloc(ctx, GG));
  DMO_TRY(slotA.alloc(ctx, n));
  DMO_TRY(slotB.alloc(ctx, n));
  DMO_TRY(crecA.alloc(ctx, n));
  DMO_TRY(crecB.alloc(ctx, n));
  DMO_TRY(alive.alloc(ctx, n));
  DMO_TRY(dom.alloc(ctx, n));
  DMO_TRY(count.alloc(ctx, 1));
  DMO_LAUNCH(peel_key_kernel, g, 256, 0, R, n, maxid, gbits, key0.p, keyA.p, keyB.p);
  DMO_TRY(prim_iota_u32(ctx, iota.p, n));
  DMO_TRY(prim_sort_pairs_u32(ctx, key0.p, keyS.p, iota.p, ord0.p, n, 0, bits));  // by objective-1 id ...
  for (int pass = 0; pass < 2; ++pass) {                                        // ... then stably by cell
    DMO_LAUNCH(gather_u32_kernel, g, 256, 0, pass == 0 ? keyA.p : keyB.p, ord0.p, n, keyT.p);
    DMO_TRY(prim_sort_pairs_u32(ctx, keyT.p, keyS.p, ord0.p, ordS.p, n, 0, 2 * gbits));
    DMO_LAUNCH(peel_gather_kernel, g, 256, 0, R, ordS.p, n, pass == 0 ? crecA.p : crecB.p, pass == 0 ? slotA.p : slotB.p);
    DMO_LAUNCH(ndg_start_kernel, (unsigned)ceil_div(GG + 1, 256), 256, 0, keyS.p, n, GG, pass == 0 ? cstartA.p : cstartB.p);
  }
  DMO_CUDA(cudaMemcpyAsync(firstA.p, cstartA.p, (size_t)GG * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  DMO_LAUNCH(fill_u8_kernel, g, 256, 0, alive.p, n, (uint8_t)1);
  DMO_CUDA(cudaMemsetAsync(count.p, 0, sizeof(unsigned long long), ctx->stream));
  unsigned long long ranked = 0, before = 0;
  double prev_front = 0.0;
  int k = 0;
  for (;; ++k) {
    DMO_LAUNCH(peel_cellmin_kernel, (unsigned)ceil_div(GG, 256), 256, 0, cstartA.p, crecA.p, GG, firstA.p, pm.p);
    DMO_LAUNCH(peel_prefix_min_kernel, G, G, G * sizeof(uint32_t), pm.p, gbits, 0);
    DMO_LAUNCH(peel_prefix_min_kernel, G, G, G * sizeof(uint32_t), pm.p, gbits, 1);
    DMO_LAUNCH(peel_flag_kernel, (unsigned)ceil_div(n, 128), 128, 0, R, n, maxid, gbits, pm.p, cstartA.p, cstartB.p, crecA.p, crecB.p,
               alive.p, dom.p);
    DMO_LAUNCH(peel_mark_kernel, g, 256, 0, n, alive.p, dom.p, k, d_rank, slotA.p, slotB.p, crecA.p, crecB.p, count.p);
    DMO_CHECK_LAUNCH();
    before = ranked;
    DMO_CUDA(cudaMemcpyAsync(&ranked, count.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    DMO_CUDA(cudaStreamSynchronize(ctx->stream));
    if ((int64_t)ranked >= keep || (int64_t)ranked >= n) break;
    // Give up when many peels are still ahead.  Fronts usually grow over the first few peels (the bench's sets: 2.4 k,
    // then 5 - 12 k points per front), so the forecast extrapolates the last two front sizes linearly, and the first
    // front only decides when it is tiny (a uniform cloud: 72 of 131 072 points).
    const double last = (double)(ranked - before);
    bool give_up;
    if (k == 0) {
      give_up = last * 64.0 < (double)keep;
    } else {
      const double rem = (double)(keep - (int64_t)ranked);
      const double grow = last > prev_front ? last - prev_front : 0.0;
      const double b = last + 0.5 * grow;
      const double ahead = grow > 0.0 ? (-b + sqrt(b * b + 2.0 * grow * rem)) / grow : rem / (last > 0.0 ? last : 1.0);
      give_up = (double)(k + 1) + ahead > (double)max_peels;
    }
    if (give_up) return DMO_OK;  // *done stays false: the chain takes over
    prev_front = last;
  }
  DMO_LAUNCH(peel_rest_kernel, g, 256, 0, n, alive.p, k + 1, d_rank);
  DMO_CHECK_LAUNCH();
  *done = true;
  return DMO_OK;
}

}  // namespace

int rank_nd_device_ex(dmo_ctx* ctx, const double* dY, int64_t n, int M, int32_t* d_rank, bool flags_only, int64_t keep = 0) {
  if (n <= 0) return DMO_OK;
  DMO_REQUIRE(M >= 1 && M <= 8, "rank_nd: M=%d out of range [1,8]", M);
  DMO_REQUIRE(n < ((int64_t)1 << 31) - 4096, "rank_nd: n too large");
  const unsigned g = (unsigned)ceil_div(n, 256);

  DevBuf<uint32_t> R;  // M x n dense integer ids (SoA)
  DevBuf<uint32_t> maxid;  // largest id per objective (= number of distinct values - 1)
  DMO_TRY(R.alloc(ctx, (size_t)M * n));
  DMO_TRY(maxid.alloc(ctx, M));
  {
    DevBuf<uint64_t> k0, k1;
    DevBuf<uint32_t> i0, i1, flag, dense;
    DMO_TRY(k0.alloc(ctx, n));
    DMO_TRY(k1.alloc(ctx, n));
    DMO_TRY(i0.alloc(ctx, n));
    DMO_TRY(i1.alloc(ctx, n));
    DMO_TRY(flag.alloc(ctx, n));
    DMO_TRY(dense.alloc(ctx, n));
    for (int j = 0; j < M; ++j) {
      DMO_LAUNCH(col_keys_kernel, g, 256, 0, dY, n, M, j, k0.p, i0.p);
      DMO_TRY(prim_sort_pairs_u64(ctx, k0.p, k1.p, i0.p, i1.p, n, 0, 64));
      DMO_LAUNCH(flag_new_u64_kernel, g, 256, 0, k1.p, n, flag.p);
      DMO_TRY(prim_inclusive_sum_u32(ctx, flag.p, dense.p, n));
      DMO_LAUNCH(scatter_dense_kernel, g, 256, 0, dense.p, i1.p, n, R.p + (size_t)j * n);
      DMO_CUDA(cudaMemcpyAsync(maxid.p + j, dense.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    DMO_CHECK_LAUNCH();
  }
  if (M == 1) {
    DMO_LAUNCH(copy_u32_to_i32_kernel, g, 256, 0, R.p, n, d_rank);
    DMO_CHECK_LAUNCH();
    return DMO_OK;
  }
  if (!flags_only && M == 3 && keep > 0 && n >= 8192 && 4 * keep <= 3 * n) {  // truncation: the best `keep` rows are enough
    bool done = false;
    DMO_TRY(rank_by_peeling(ctx, R.p, maxid.p, n, keep, d_rank, &done));
    if (done) return DMO_OK;
  }

  // lexicographic order of the id vectors: LSD passes, least significant objective first
  const int bits = bits_for(n);
  DevBuf<uint32_t> permA, permB, keyA, keyB;
  DMO_TRY(permA.alloc(ctx, n));
  DMO_TRY(permB.alloc(ctx, n));
  DMO_TRY(keyA.alloc(ctx, n));
  DMO_TRY(keyB.alloc(ctx, n));
  DMO_TRY(prim_iota_u32(ctx, permA.p, n));
  uint32_t* pin = permA.p;
  uint32_t* pout = permB.p;
  // The chain kernel for two and three objectives takes the segmented order (RankSeg): key = (segment of objective 1,
  // objective 2, ..., objective M, objective 1).  Everything else keeps the plain lexicographic order.
  const int64_t nblocks_est = ceil_div(n, RANK_T);
  int segbits = bits - 10;  // segments of 1024 dense ids of objective 1 (8 blocks), at most 128 segments
  if (segbits > 7) segbits = 7;
  if (const char* e = getenv("DMO_RANK_SEGBITS")) segbits = atoi(e);
  if (segbits < 1) segbits = 1;
  const bool use_seg = !flags_only && M <= 3 && nblocks_est >= 16 && nblocks_est <= RANK_SEG_MAXT && bits > segbits + 7 &&
                       getenv("DMO_RANK_NOSEG") == nullptr;
  const int sshift = bits - segbits;
  auto sort_pass = [&](const uint32_t* col, int shift, int nbits) -> int {
    if (shift == 0) {
      DMO_LAUNCH(gather_u32_kernel, g, 256, 0, col, pin, n, keyA.p);
    } else {
      DMO_LAUNCH(seg_key_kernel, g, 256, 0, col, pin, n, shift, keyA.p);
    }
    DMO_TRY(prim_sort_pairs_u32(ctx, keyA.p, keyB.p, pin, pout, n, 0, nbits));
    uint32_t* t = pin;
    pin = pout;
    pout = t;
    return DMO_OK;
  };
  if (use_seg) {
    DMO_TRY(sort_pass(R.p, 0, bits));  // least significant: objective 1 itself
    for (int j = M - 1; j >= 1; --j) DMO_TRY(sort_pass(R.p + (size_t)j * n, 0, bits));
    DMO_TRY(sort_pass(R.p, sshift, bits - sshift));  // most significant: the segment
  } else {
    for (int j = M - 1; j >= 0; --j) DMO_TRY(sort_pass(R.p + (size_t)j * n, 0, bits));
  }
  const uint32_t* perm = pin;

  // group ids (identical vectors share one)
  DevBuf<uint32_t> gid;
  DMO_TRY(gid.alloc(ctx, n));
  DMO_LAUNCH(flag_new_vec_kernel, g, 256, 0, R.p, perm, n, M, keyA.p);
  DMO_TRY(prim_inclusive_sum_u32(ctx, keyA.p, gid.p, n));

  const int W = 4 * ((M + 1 + 3) / 4);
  const int64_t nblocks = ceil_div(n, RANK_T);
  const int64_t npad = nblocks * RANK_T;
  DevBuf<uint32_t> rec;
  DevBuf<int> rankS, sync;
  DMO_TRY(rec.alloc(ctx, (size_t)npad * W));
  DMO_TRY(rankS.alloc(ctx, npad));
  DMO_TRY(sync.alloc(ctx, nblocks + 2));
  DMO_CUDA(cudaMemsetAsync(sync.p, 0, (nblocks + 2) * sizeof(int), ctx->stream));
  DMO_LAUNCH(build_records_kernel, (unsigned)ceil_div(npad, 256), 256, 0, R.p, perm, gid.p, n, npad, M, W, rec.p);
  int* ticket = sync.p + nblocks;
  int* errflag = sync.p + nblocks + 1;
  if (flags_only && M <= 3 && n >= 8192 && getenv("DMO_ND_BRUTE") == nullptr) {
    DMO_TRY(nd_flags_grid(ctx, rec.p, n, npad, M, maxid.p, rankS.p));
    DMO_LAUNCH(scatter_rank_kernel, g, 256, 0, rankS.p, perm, n, d_rank);
    DMO_CHECK_LAUNCH();
    return DMO_OK;
  }
  if (flags_only) {  // d_rank receives 0 for non-dominated points and 1 otherwise
    switch (M) {
      case 2: DMO_TRY(launch_nd_flags<2>(ctx, rec.p, (int)nblocks, rankS.p)); break;
      case 3: DMO_TRY(launch_nd_flags<3>(ctx, rec.p, (int)nblocks, rankS.p)); break;
      case 4: DMO_TRY(launch_nd_flags<4>(ctx, rec.p, (int)nblocks, rankS.p)); break;
      case 5: DMO_TRY(launch_nd_flags<5>(ctx, rec.p, (int)nblocks, rankS.p)); break;
      case 6: DMO_TRY(launch_nd_flags<6>(ctx, rec.p, (int)nblocks, rankS.p)); break;
      case 7: DMO_TRY(launch_nd_flags<7>(ctx, rec.p, (int)nblocks, rankS.p)); break;
      default: DMO_TRY(launch_nd_flags<8>(ctx, rec.p, (int)nblocks, rankS.p)); break;
    }
    DMO_LAUNCH(scatter_rank_kernel, g, 256, 0, rankS.p, perm, n, d_rank);
    DMO_CHECK_LAUNCH();
    return DMO_OK;
  }
  RankSeg sg;
  DevBuf<uint32_t> c1rec, seg_start;
  DevBuf<uint16_t> tile_q;
  DevBuf<unsigned long long> stair;
  if (use_seg) {
    DMO_TRY(stair.alloc(ctx, npad));
    DMO_CUDA(cudaMemsetAsync(stair.p, 0, (size_t)npad * sizeof(unsigned long long), ctx->stream));
    const int nseg = (int)(((uint32_t)(n - 1)) >> sshift) + 1;
    int qshift = bits - 8;
    if (qshift < 0) qshift = 0;
    DMO_TRY(c1rec.alloc(ctx, npad));
    DMO_TRY(seg_start.alloc(ctx, nseg + 1));
    DMO_TRY(tile_q.alloc(ctx, nblocks));
    DMO_LAUNCH(seg_c1_kernel, (unsigned)ceil_div(npad, 256), 256, 0, R.p, perm, n, npad, c1rec.p);
    DMO_LAUNCH(seg_start_kernel, (unsigned)ceil_div(nseg + 1, 128), 128, 0, c1rec.p, n, sshift, nseg, seg_start.p);
    DMO_LAUNCH(seg_tile_band_kernel, (unsigned)ceil_div(nblocks * 32, 256), 256, 0, rec.p, (int)nblocks, RANK_T, qshift, tile_q.p);
    DMO_CHECK_LAUNCH();
    sg.c1rec = c1rec.p;
    sg.seg_start = seg_start.p;
    sg.tile_q = tile_q.p;
    sg.stair = stair.p;
    sg.sshift = sshift;
    sg.qshift = qshift;
    if (M == 2) {
      DMO_TRY((launch_chain<2, true>(ctx, rec.p, (int)nblocks, rankS.p, ticket, errflag, sg)));
    } else {
      DMO_TRY((launch_chain<3, true>(ctx, rec.p, (int)nblocks, rankS.p, ticket, errflag, sg)));
    }
  } else {
    switch (M) {
      case 2: DMO_TRY((launch_chain<2, false>(ctx, rec.p, (int)nblocks, rankS.p, ticket, errflag, sg))); break;
      case 3: DMO_TRY((launch_chain<3, false>(ctx, rec.p, (int)nblocks, rankS.p, ticket, errflag, sg))); break;
      case 4: DMO_TRY((launch_chain<4, false>(ctx, rec.p, (int)nblocks, rankS.p, ticket, errflag, sg))); break;
      case 5: DMO_TRY((launch_chain<5, false>(ctx, rec.p, (int)nblocks, rankS.p, ticket, errflag, sg))); break;
      case 6: DMO_TRY((launch_chain<6, false>(ctx, rec.p, (int)nblocks, rankS.p, ticket, errflag, sg))); break;
      case 7: DMO_TRY((launch_chain<7, false>(ctx, rec.p, (int)nblocks, rankS.p, ticket, errflag, sg))); break;
      default: DMO_TRY((launch_chain<8, false>(ctx, rec.p, (int)nblocks, rankS.p, ticket, errflag, sg))); break;
    }
  }
  DMO_LAUNCH(scatter_rank_kernel, g, 256, 0, rankS.p, perm, n, d_rank);
  DMO_CHECK_LAUNCH();
  int herr = 0;
  DMO_CUDA(cudaMemcpyAsync(&herr, errflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  DMO_CUDA(cudaStreamSynchronize(ctx->stream));
  if (herr) return dmo_fail(ctx, DMO_ERR_INTERNAL, "rank_nd: chain kernel watchdog tripped");
  return DMO_OK;
}

int rank_nd_device(dmo_ctx* ctx, const double* dY, int64_t n, int M, int32_t* d_rank) {
  return rank_nd_device_ex(ctx, dY, n, M, d_rank, false);
}
// ranks that are exact for (at least) the `keep` best rows; the other rows share one larger value (see rank_by_peeling)
int rank_nd_device_keep(dmo_ctx* ctx, const double* dY, int64_t n, int M, int64_t keep, int32_t* d_rank) {
  return rank_nd_device_ex(ctx, dY, n, M, d_rank, false, keep);
}
int nondominated_flags_device(dmo_ctx* ctx, const double* dY, int64_t n, int M, int32_t* d_flag01) {
  return rank_nd_device_ex(ctx, dY, n, M, d_flag01, true);
}

extern "C" int dmo_rank_nd(dmo_ctx* ctx, const double* Y, int64_t n, int M, int32_t* rank) {
  if (!ctx) return DMO_ERR_ARG;
  DMO_CUDA(cudaSetDevice(ctx->device));
  DMO_REQUIRE(n >= 0 && M >= 1, "rank_nd: bad shape n=%lld M=%d", (long long)n, M);
  if (n == 0) return DMO_OK;
  DMO_REQUIRE(Y && rank, "rank_nd: null pointer");
  In<double> y;
  Out<int32_t> r;
  DMO_TRY(y.init(ctx, Y, (size_t)n * M));
  DMO_TRY(r.init(ctx, rank, (size_t)n));
  DMO_TRY(rank_nd_device(ctx, y.d, n, M, r.d));
  DMO_TRY(r.finish(ctx));
  DMO_CUDA(cudaStreamSynchronize(ctx->stream));
  return DMO_OK;
}
