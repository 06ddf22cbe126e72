// Patch-graph index structure, built on the device in ONE cooperative launch.
//
// The reference rebuilds the same information three times per update() on different paths:
//   * fastba.neighbors   ba.cpp:59-97      torch::_unique on GPU, D2H, per-group CPU stable_sort, H2D
//   * SoftAgg            blocks.py:41      torch.unique(return_inverse) x2 (host sync each)
//   * cuda_ba            ba_cuda.cu:447    torch::_unique(kk) (host sync)
// Here: a stable LSD radix sort of edge ids by (key_a, key_b, sec) over only the bits the value
// ranges need (ranges are found on the device), followed by boundary flagging and a scan that
// yields a CSR (order / group_start), the dense group id per edge and the group keys.  All
// phases run inside one persistent kernel separated by a software grid barrier; there are no
// floating-point or order-dependent atomics, so the result is bit-reproducible.
#include "common.cuh"

namespace dpvo {

constexpr int G_THREADS = 256;
constexpr int G_ITEMS = 8;
constexpr int G_TILE = G_THREADS * G_ITEMS;   // 2048 edge slots per tile
constexpr int G_WARPS = G_THREADS / 32;
constexpr int G_BINS = 256;

struct GroupHeader {            // lives at the start of the workspace
  unsigned barrier;             // must be 0 at launch (memset by the host wrapper)
  int npass[3];                 // radix passes for sec / key_b / key_a
  long long amin, amax, bmin, bmax, smin, smax;
};

struct GroupArgs {
  const int64_t* ka; const int64_t* kb; const int64_t* sec;
  int64_t E; int ntiles;
  int32_t* order; int32_t* group_of; int32_t* group_start;
  int64_t* group_key_a; int64_t* group_key_b; int32_t* n_groups;
  GroupHeader* hdr; int32_t* buf0; int32_t* buf1; unsigned* hist; int32_t* tile_sums;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    while (ld_acquire_u32(ctr) < epoch) { }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ int bits_needed(long long range) {   // range >= 0
  int b = 0;
  while (range > 0) { ++b; range >>= 1; }
  return b;
}

// exclusive scan of one int per thread across the block; returns the exclusive prefix and the
// block total
__device__ __forceinline__ int block_excl_scan(int v, int* smem_warp /*[G_WARPS+1]*/, int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();
  if (lane == 31) smem_warp[w] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < G_WARPS; ++i) { int t = smem_warp[i]; smem_warp[i] = run; run += t; }
    smem_warp[G_WARPS] = run;
  }
  __syncthreads();
  total = smem_warp[G_WARPS];
  return smem_warp[w] + inc - v;
}

__global__ void __launch_bounds__(G_THREADS)
group_edges_kernel(const GroupArgs a) {
  __shared__ unsigned warp_hist[G_WARPS][G_BINS];
  __shared__ int scan_tmp[G_WARPS + 1];
  __shared__ long long red[6];
  unsigned epoch = 0;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  GroupHeader* H = a.hdr;
  const int64_t E = a.E;

  // ---- phase 0a: initialise the header
  if (blockIdx.x == 0 && tid == 0) {
    H->amin = H->bmin = H->smin = 0x7fffffffffffffffLL;
    H->amax = H->bmax = H->smax = -0x7fffffffffffffffLL - 1;
  }
  grid_barrier(&H->barrier, epoch);

  // ---- phase 0b: value ranges, identity permutation
  {
    long long mn[3] = {0x7fffffffffffffffLL, 0x7fffffffffffffffLL, 0x7fffffffffffffffLL};
    long long mx[3] = {-0x7fffffffffffffffLL - 1, -0x7fffffffffffffffLL - 1, -0x7fffffffffffffffLL - 1};
    for (int64_t i = (int64_t)blockIdx.x * G_THREADS + tid; i < E; i += (int64_t)gridDim.x * G_THREADS) {
      a.buf0[i] = (int32_t)i;
      const long long va = a.ka[i];
      mn[0] = min(mn[0], va); mx[0] = max(mx[0], va);
      if (a.kb) { const long long vb = a.kb[i]; mn[1] = min(mn[1], vb); mx[1] = max(mx[1], vb); }
      if (a.sec) { const long long vs = a.sec[i]; mn[2] = min(mn[2], vs); mx[2] = max(mx[2], vs); }
    }
    if (tid < 6) red[tid] = (tid < 3) ? 0x7fffffffffffffffLL : (-0x7fffffffffffffffLL - 1);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn[k] = min(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
        mx[k] = max(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
      }
      if (lane == 0) { atomicMin(&red[k], mn[k]); atomicMax(&red[3 + k], mx[k]); }
    }
    __syncthreads();
    if (tid == 0) {
      atomicMin(&H->amin, red[0]); atomicMax(&H->amax, red[3]);
      if (a.kb) { atomicMin(&H->bmin, red[1]); atomicMax(&H->bmax, red[4]); }
      if (a.sec) { atomicMin(&H->smin, red[2]); atomicMax(&H->smax, red[5]); }
    }
  }
  grid_barrier(&H->barrier, epoch);

  const long long amin = H->amin, bmin = a.kb ? H->bmin : 0, smin = a.sec ? H->smin : 0;
  int npass[3];
  npass[0] = a.sec ? (bits_needed(H->smax - smin) + 7) / 8 : 0;
  npass[1] = a.kb ? (bits_needed(H->bmax - bmin) + 7) / 8 : 0;
  npass[2] = (bits_needed(H->amax - amin) + 7) / 8;

  int32_t* src = a.buf0;
  int32_t* dst = a.buf1;

  // ---- radix passes: sec (least significant field) -> key_b -> key_a
  for (int field = 0; field < 3; ++field) {
    const int64_t* vals = field == 0 ? a.sec : (field == 1 ? a.kb : a.ka);
    const long long vmin = field == 0 ? smin : (field == 1 ? bmin : amin);
    for (int pass = 0; pass < npass[field]; ++pass) {
      const int shift = 8 * pass;

      // count: stable per-warp digit histograms of every tile this block owns
      for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        for (int i = tid; i < G_WARPS * G_BINS; i += G_THREADS) (&warp_hist[0][0])[i] = 0u;
        __syncthreads();
        const int64_t base = (int64_t)t * G_TILE + w * (G_TILE / G_WARPS);
        for (int c = 0; c < G_TILE / G_WARPS / 32; ++c) {
          const int64_t i = base + c * 32 + lane;
          unsigned d = 0xffffffffu;
          if (i < E) d = (unsigned)(((unsigned long long)(vals[src[i]] - vmin) >> shift) & 255ull);
          const unsigned peers = __match_any_sync(0xffffffffu, d);
          if (d != 0xffffffffu && lane == __ffs(peers) - 1) warp_hist[w][d] += __popc(peers);
          __syncwarp();
        }
        __syncthreads();
        unsigned s = 0;
#pragma unroll
        for (int k = 0; k < G_WARPS; ++k) s += warp_hist[k][tid];
        a.hist[(size_t)tid * a.ntiles + t] = s;
        __syncthreads();
      }
      grid_barrier(&H->barrier, epoch);

      // scan (block 0): exclusive prefix over hist in bin-major order
      if (blockIdx.x == 0) {
        unsigned* h = a.hist + (size_t)tid * a.ntiles;
        int tot = 0;
        for (int t = 0; t < a.ntiles; ++t) tot += (int)h[t];
        int all;
        int run = block_excl_scan(tot, scan_tmp, all);
        for (int t = 0; t < a.ntiles; ++t) { const unsigned v = h[t]; h[t] = (unsigned)run; run += (int)v; }
      }
      grid_barrier(&H->barrier, epoch);

      // scatter
      for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        for (int i = tid; i < G_WARPS * G_BINS; i += G_THREADS) (&warp_hist[0][0])[i] = 0u;
        __syncthreads();
        const int64_t base = (int64_t)t * G_TILE + w * (G_TILE / G_WARPS);
        for (int c = 0; c < G_TILE / G_WARPS / 32; ++c) {
          const int64_t i = base + c * 32 + lane;
          unsigned d = 0xffffffffu;
          if (i < E) d = (unsigned)(((unsigned long long)(vals[src[i]] - vmin) >> shift) & 255ull);
          const unsigned peers = __match_any_sync(0xffffffffu, d);
          if (d != 0xffffffffu && lane == __ffs(peers) - 1) warp_hist[w][d] += __popc(peers);
          __syncwarp();
        }
        __syncthreads();
        {   // per-bin exclusive prefix over the warps, offset by the global position of (bin, tile)
          unsigned run = a.hist[(size_t)tid * a.ntiles + t];
#pragma unroll
          for (int k = 0; k < G_WARPS; ++k) { const unsigned v = warp_hist[k][tid]; warp_hist[k][tid] = run; run += v; }
        }
        __syncthreads();
        for (int c = 0; c < G_TILE / G_WARPS / 32; ++c) {
          const int64_t i = base + c * 32 + lane;
          unsigned d = 0xffffffffu;
          int32_t id = 0;
          if (i < E) { id = src[i]; d = (unsigned)(((unsigned long long)(vals[id] - vmin) >> shift) & 255ull); }
          const unsigned peers = __match_any_sync(0xffffffffu, d);
          const int leader = __ffs(peers) - 1;
          unsigned pos = 0;
          if (d != 0xffffffffu && lane == leader) { pos = warp_hist[w][d]; warp_hist[w][d] = pos + __popc(peers); }
          pos = __shfl_sync(0xffffffffu, pos, leader);
          if (d != 0xffffffffu) dst[pos + __popc(peers & ((1u << lane) - 1u))] = id;
          __syncwarp();
        }
        __syncthreads();
      }
      grid_barrier(&H->barrier, epoch);
      int32_t* tmp = src; src = dst; dst = tmp;
    }
  }

  // ---- boundaries: flag[i] = first position of a new (key_a, key_b) group
  auto is_head = [&](int64_t i) -> int {
    if (i == 0) return 1;
    const int32_t e = src[i], p = src[i - 1];
    if (a.ka[e] != a.ka[p]) return 1;
    if (a.kb && a.kb[e] != a.kb[p]) return 1;
    return 0;
  };
  for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
    int cnt = 0;
    const int64_t base = (int64_t)t * G_TILE + (int64_t)tid * G_ITEMS;
#pragma unroll
    for (int k = 0; k < G_ITEMS; ++k) if (base + k < E) cnt += is_head(base + k);
    int total;
    block_excl_scan(cnt, scan_tmp, total);
    if (tid == 0) a.tile_sums[t] = total;
    __syncthreads();
  }
  grid_barrier(&H->barrier, epoch);
  if (blockIdx.x == 0) {
    // exclusive scan of tile_sums, chunked through the block
    int carry = 0;
    for (int t0 = 0; t0 < a.ntiles; t0 += G_THREADS) {
      const int t = t0 + tid;
      const int v = (t < a.ntiles) ? a.tile_sums[t] : 0;
      int total;
      const int ex = block_excl_scan(v, scan_tmp, total);
      if (t < a.ntiles) a.tile_sums[t] = carry + ex;
      carry += total;
      __syncthreads();
    }
    if (tid == 0) { *a.n_groups = carry; a.group_start[carry] = (int32_t)E; }
  }
  grid_barrier(&H->barrier, epoch);
  for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
    int flags[G_ITEMS];
    int cnt = 0;
    const int64_t base = (int64_t)t * G_TILE + (int64_t)tid * G_ITEMS;
#pragma unroll
    for (int k = 0; k < G_ITEMS; ++k) { flags[k] = (base + k < E) ? is_head(base + k) : 0; cnt += flags[k]; }
    int total;
    int gid = a.tile_sums[t] + block_excl_scan(cnt, scan_tmp, total) - 1;
#pragma unroll
    for (int k = 0; k < G_ITEMS; ++k) {
      const int64_t i = base + k;
      if (i < E) {
        const int32_t e = src[i];
        gid += flags[k];
        a.group_of[e] = gid;
        a.order[i] = e;
        if (flags[k]) {
          a.group_start[gid] = (int32_t)i;
          a.group_key_a[gid] = a.ka[e];
          if (a.group_key_b) a.group_key_b[gid] = a.kb ? a.kb[e] : 0;
        }
      }
    }
    __syncthreads();
  }
}

// neighbours inside a group in sorted order (ba.cpp:88-94)
__global__ void neighbors_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ group_of,
                                 int64_t E, int64_t* __restrict__ ix, int64_t* __restrict__ jx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t e = order[i];
    const int32_t g = group_of[e];
    int64_t prev = -1, next = -1;
    if (i > 0) { const int32_t p = order[i - 1]; if (group_of[p] == g) prev = p; }
    if (i + 1 < E) { const int32_t n = order[i + 1]; if (group_of[n] == g) next = n; }
    ix[e] = prev; jx[e] = next;
  }
}

static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

static int64_t group_ws_bytes(int64_t E) {
  const int64_t ntiles = (E + G_TILE - 1) / G_TILE;
  return align256(sizeof(GroupHeader)) + 2 * align256(E * 4) + align256(ntiles * G_BINS * 4) + align256((ntiles + 1) * 4);
}

static int group_launch(const int64_t* ka, const int64_t* kb, const int64_t* sec, int64_t E,
                        int32_t* order, int32_t* group_of, int32_t* group_start,
                        int64_t* gka, int64_t* gkb, int32_t* n_groups,
                        void* ws, int64_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < group_ws_bytes(E)) {
    set_error("group_edges: workspace %lld B < required %lld B", (long long)ws_bytes, (long long)group_ws_bytes(E));
    return DPVO_ERR_WORKSPACE;
  }
  if (E >= (1ll << 31)) { set_error("group_edges: E too large"); return DPVO_ERR_UNSUPPORTED; }
  GroupArgs a;
  a.ka = ka; a.kb = kb; a.sec = sec; a.E = E; a.ntiles = (int)((E + G_TILE - 1) / G_TILE);
  a.order = order; a.group_of = group_of; a.group_start = group_start;
  a.group_key_a = gka; a.group_key_b = gkb; a.n_groups = n_groups;
  char* p = (char*)ws;
  a.hdr = (GroupHeader*)p; p += align256(sizeof(GroupHeader));
  a.buf0 = (int32_t*)p; p += align256(E * 4);
  a.buf1 = (int32_t*)p; p += align256(E * 4);
  a.hist = (unsigned*)p; p += align256((int64_t)a.ntiles * G_BINS * 4);
  a.tile_sums = (int32_t*)p;
  int rc = check_cuda(cudaMemsetAsync(a.hdr, 0, sizeof(GroupHeader), st), "group_edges: memset");
  if (rc) return rc;
  if (E == 0) {
    rc = check_cuda(cudaMemsetAsync(n_groups, 0, 4, st), "group_edges: memset");
    if (rc) return rc;
    return check_cuda(cudaMemsetAsync(group_start, 0, 4, st), "group_edges: memset");
  }
  static int max_blocks = 0;
  if (max_blocks == 0) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, group_edges_kernel, G_THREADS, 0);
    max_blocks = std::max(1, per_sm) * sm_count();
  }
  const int grid = std::min(a.ntiles, max_blocks);
  void* kargs[] = {(void*)&a};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)group_edges_kernel, dim3(grid), dim3(G_THREADS), kargs, 0, st);
  if (e != cudaSuccess) return check_cuda(e, "group_edges_kernel");
  DPVO_LAUNCH_CHECK("group_edges_kernel");
  return DPVO_OK;
}

}  // namespace dpvo

using namespace dpvo;

extern "C" int64_t dpvo_group_workspace_bytes(int64_t E) { return group_ws_bytes(E < 0 ? 0 : E); }

extern "C" int dpvo_group_edges(const int64_t* key_a, const int64_t* key_b, const int64_t* sec, int64_t E,
                                int32_t* order, int32_t* group_of, int32_t* group_start,
                                int64_t* group_key_a, int64_t* group_key_b, int32_t* n_groups,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  DPVO_REQUIRE(E >= 0, "group_edges: negative E");
  DPVO_REQUIRE(order && group_of && group_start && group_key_a && n_groups && workspace, "group_edges: null pointer");
  DPVO_REQUIRE(E == 0 || key_a, "group_edges: null key");
  return group_launch(key_a, key_b, sec, E, order, group_of, group_start, group_key_a, group_key_b, n_groups,
                      workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int64_t dpvo_neighbors_workspace_bytes(int64_t E) {
  if (E < 0) E = 0;
  // order, group_of, group_start(+1), group_key, n_groups + the grouping workspace
  return group_ws_bytes(E) + 2 * align256(E * 4) + align256((E + 1) * 4) + align256(E * 8) + 256;
}

extern "C" int dpvo_neighbors(const int64_t* ii, const int64_t* jj, int64_t E,
                              int64_t* ix, int64_t* jx,
                              void* workspace, int64_t workspace_bytes, void* stream) {
  DPVO_REQUIRE(E >= 0, "neighbors: negative E");
  if (E == 0) return DPVO_OK;
  DPVO_REQUIRE(ii && jj && ix && jx && workspace, "neighbors: null pointer");
  if (workspace_bytes < dpvo_neighbors_workspace_bytes(E)) {
    set_error("neighbors: workspace too small");
    return DPVO_ERR_WORKSPACE;
  }
  char* p = (char*)workspace;
  int32_t* order = (int32_t*)p; p += align256(E * 4);
  int32_t* group_of = (int32_t*)p; p += align256(E * 4);
  int32_t* group_start = (int32_t*)p; p += align256((E + 1) * 4);
  int64_t* gkey = (int64_t*)p; p += align256(E * 8);
  int32_t* ng = (int32_t*)p; p += 256;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = group_launch(ii, nullptr, jj, E, order, group_of, group_start, gkey, nullptr, ng, p,
                        workspace_bytes - (p - (char*)workspace), st);
  if (rc) return rc;
  const int threads = 256;
  const unsigned blocks = (unsigned)std::min<int64_t>((E + threads - 1) / threads, (int64_t)sm_count() * 8);
  neighbors_kernel<<<blocks, threads, 0, st>>>(order, group_of, E, ix, jx);
  DPVO_LAUNCH_CHECK("neighbors_kernel");
  return DPVO_OK;
}

extern "C" int dpvo_neighbors_from_groups(const int32_t* order, const int32_t* group_of, int64_t E,
                                          int64_t* ix, int64_t* jx, void* stream) {
  DPVO_REQUIRE(E >= 0, "neighbors_from_groups: negative E");
  if (E == 0) return DPVO_OK;
  DPVO_REQUIRE(order && group_of && ix && jx, "neighbors_from_groups: null pointer");
  const int threads = 256;
  const unsigned blocks = (unsigned)std::min<int64_t>((E + threads - 1) / threads, (int64_t)sm_count() * 8);
  neighbors_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(order, group_of, E, ix, jx);
  DPVO_LAUNCH_CHECK("neighbors_kernel");
  return DPVO_OK;
}
