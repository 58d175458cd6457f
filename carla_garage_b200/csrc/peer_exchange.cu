// Data-parallel gradient exchange over NVLink peer memory, fused with the optimizer (one process per GPU).
//
// Reference: train.py:516 wraps the model in DistributedDataParallel (NCCL all-reduce of every gradient bucket) and
// train.py:527-531 shards AdamW over the ranks (ZeroRedundancyOptimizer: each rank steps 1/W of the parameters, then
// broadcasts them).  Here that is ONE kernel per step and rank:
//
//   for the shard of the flat buffers this rank owns:
//     g     = sum over ranks of grad_r[i]            (W-1 peer loads over NVLink: the reduce-scatter)
//     p, m, v, vmax = AdamW(amsgrad)(p, g / W)       (the ZeRO-1 optimizer shard)
//     param_r[i] = p for every rank r                (W-1 peer stores over NVLink: the all-gather)
//
// bracketed by two flag barriers (every rank's gradients are complete before anyone reads them; every rank's pushes
// have landed before anyone rebuilds weight packs or zeroes gradients).  The buffers are cudaMalloc'ed here and mapped
// into the peers with CUDA IPC; all of it is plain kernels, so the whole training step (forward, backward, exchange,
// optimizer) stays inside ONE CUDA graph at any world size — NCCL collectives had to stay outside the capture
//  and cost 2.3 ms exposed per step at 8 GPUs in round 1.
#include "../../include/tfpp.h"
#include <string.h>

#include "common.cuh"

namespace {

struct Peers {
  int world, rank;
  float* grad[TFPP_MAX_PEERS];
  float* param[TFPP_MAX_PEERS];
  unsigned* flags[TFPP_MAX_PEERS];
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Flag layout in every rank's array (uint32): [slot 0: world arrival flags][slot 1: world arrival flags] at stride
// TFPP_MAX_PEERS, then two local epoch counters.  Barrier `slot`: bump the local epoch e, write e into entry `rank` of
// every peer's slot, wait until all `world` entries of the own slot reached e.  Bounded spin (~20 s): a lost peer
// becomes a CUDA error instead of a hung GPU.
__global__ void peer_barrier_kernel(const Peers p, int slot) {
  unsigned* mine = p.flags[p.rank];
  __shared__ unsigned epoch;
  if (threadIdx.x == 0) epoch = mine[2 * TFPP_MAX_PEERS + slot] + 1u;
  __syncthreads();
  const unsigned e = epoch;
  if (threadIdx.x < p.world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + slot * TFPP_MAX_PEERS + p.rank, e);
    const unsigned* w = mine + slot * TFPP_MAX_PEERS + threadIdx.x;
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_sys(w) - e) < 0) {
      if (clock64() - t0 > 40000000000ll) __trap();
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) mine[2 * TFPP_MAX_PEERS + slot] = e;
}

__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, float& vmax, float lr, float beta1,
                                           float beta2, float eps, float wd, float bc1, float bc2_sqrt) {
  const float pv = p * (1.f - lr * wd);
  m = beta1 * m + (1.f - beta1) * g;
  v = beta2 * v + (1.f - beta2) * g * g;
  vmax = fmaxf(vmax, v);
  p = pv - (lr / bc1) * (m / (sqrtf(vmax) / bc2_sqrt + eps));
}

template <int W>
__global__ void __launch_bounds__(256) peer_adamw_kernel(const Peers pr, float* __restrict__ em, float* __restrict__ ev,
                                                         float* __restrict__ evmax, long long lo4, long long hi4,
                                                         float beta1, float beta2, float eps, float wd,
                                                         const float* __restrict__ dev_state,
                                                         const unsigned char* __restrict__ flags) {
  const float lr = dev_state[1], bc1 = dev_state[2], bc2s = dev_state[3];
  const float inv_w = 1.f / static_cast<float>(W);
  const int me = pr.rank;
  for (long long i4 = lo4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i4 < hi4;
       i4 += static_cast<long long>(gridDim.x) * blockDim.x) {
    // reduce-scatter: all W gradient loads are in flight together (peer loads bypass L1: the same addresses carry new
    // values every step)
    float4 g[W];
#pragma unroll
    for (int r = 0; r < W; ++r) g[r] = __ldcv(reinterpret_cast<const float4*>(pr.grad[r]) + i4);
    float4 s = g[0];
#pragma unroll
    for (int r = 1; r < W; ++r) {
      s.x += g[r].x; s.y += g[r].y; s.z += g[r].z; s.w += g[r].w;
    }
    s.x *= inv_w; s.y *= inv_w; s.z *= inv_w; s.w *= inv_w;
    unsigned fl = flags ? *reinterpret_cast<const unsigned*>(flags + 4 * i4) : 0u;
    if ((fl & 0x02020202u) == 0x02020202u) continue;  // frozen on every rank: nothing changes anywhere
    float4 pv = reinterpret_cast<float4*>(pr.param[me])[i4];
    float4 mv = reinterpret_cast<float4*>(em)[i4], vv = reinterpret_cast<float4*>(ev)[i4];
    float4 xv = reinterpret_cast<float4*>(evmax)[i4];
    if (!(fl & 0x00000002u)) adamw_elem(pv.x, s.x, mv.x, vv.x, xv.x, lr, beta1, beta2, eps, (fl & 0x00000001u) ? 0.f : wd, bc1, bc2s);
    if (!(fl & 0x00000200u)) adamw_elem(pv.y, s.y, mv.y, vv.y, xv.y, lr, beta1, beta2, eps, (fl & 0x00000100u) ? 0.f : wd, bc1, bc2s);
    if (!(fl & 0x00020000u)) adamw_elem(pv.z, s.z, mv.z, vv.z, xv.z, lr, beta1, beta2, eps, (fl & 0x00010000u) ? 0.f : wd, bc1, bc2s);
    if (!(fl & 0x02000000u)) adamw_elem(pv.w, s.w, mv.w, vv.w, xv.w, lr, beta1, beta2, eps, (fl & 0x01000000u) ? 0.f : wd, bc1, bc2s);
    reinterpret_cast<float4*>(em)[i4] = mv;
    reinterpret_cast<float4*>(ev)[i4] = vv;
    reinterpret_cast<float4*>(evmax)[i4] = xv;
    // all-gather: the updated shard goes to every replica (the local copy included)
#pragma unroll
    for (int r = 0; r < W; ++r) reinterpret_cast<float4*>(pr.param[r])[i4] = pv;
  }
  __threadfence_system();  // the pushes have landed before this thread's block can count as done
}

__global__ void peer_tick_kernel(float* dev_state, float beta1, float beta2) {
  const float t = dev_state[0] + 1.f;
  dev_state[0] = t;
  dev_state[2] = 1.f - powf(beta1, t);
  dev_state[3] = sqrtf(1.f - powf(beta2, t));
}

}  // namespace

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      tfpp_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_));        \
      return TFPP_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

extern "C" int tfpp_peer_alloc(long long bytes, void** ptr, void* handle64) {
  TFPP_CHECK_ARG(bytes > 0 && ptr != nullptr && handle64 != nullptr, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, static_cast<size_t>(bytes)));
  CUDA_TRY(cudaMemset(p, 0, static_cast<size_t>(bytes)));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  *ptr = p;
  return TFPP_OK;
}

extern "C" int tfpp_peer_open(const void* handle64, void** ptr) {
  TFPP_CHECK_ARG(ptr != nullptr && handle64 != nullptr, "bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return TFPP_OK;
}

extern "C" int tfpp_peer_close(void* ptr) {
  CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return TFPP_OK;
}

extern "C" int tfpp_peer_free(void* ptr) {
  CUDA_TRY(cudaFree(ptr));
  return TFPP_OK;
}

static int fill_peers(const tfpp_peer_step_args* a, Peers* p) {
  TFPP_CHECK_ARG(a != nullptr && a->world >= 1 && a->world <= TFPP_MAX_PEERS && a->rank >= 0 && a->rank < a->world,
                 "world must be 1..8 and rank < world");
  p->world = a->world;
  p->rank = a->rank;
  for (int r = 0; r < TFPP_MAX_PEERS; ++r) {
    p->grad[r] = r < a->world ? a->grad[r] : nullptr;
    p->param[r] = r < a->world ? a->param[r] : nullptr;
    p->flags[r] = r < a->world ? a->flags[r] : nullptr;
    if (r < a->world) TFPP_CHECK_ARG(a->flags[r] != nullptr, "null peer flag array");
  }
  return TFPP_OK;
}

extern "C" int tfpp_peer_barrier(const tfpp_peer_step_args* a, int slot, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Peers p;
  int rc = fill_peers(a, &p);
  if (rc) return rc;
  TFPP_CHECK_ARG(slot == 0 || slot == 1, "slot is 0 or 1");
  peer_barrier_kernel<<<1, 32, 0, stream>>>(p, slot);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_peer_adamw_step(const tfpp_peer_step_args* a, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Peers p;
  int rc = fill_peers(a, &p);
  if (rc) return rc;
  TFPP_CHECK_ARG(a->n > 0 && a->n % 4 == 0, "n must be a positive multiple of 4");
  TFPP_CHECK_ARG(a->dev_state != nullptr && a->exp_avg != nullptr && a->exp_avg_sq != nullptr && a->max_exp_avg_sq != nullptr,
                 "optimizer state is required");
  for (int r = 0; r < a->world; ++r)
    TFPP_CHECK_ARG(a->grad[r] != nullptr && a->param[r] != nullptr && (reinterpret_cast<uintptr_t>(a->grad[r]) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a->param[r]) & 15) == 0,
                   "peer buffers must be non-null and 16-byte aligned");
  const long long n4 = a->n / 4;
  const long long chunk = ceil_div_ll(n4, a->world);
  const long long lo4 = chunk * a->rank, hi4 = (lo4 + chunk < n4) ? lo4 + chunk : n4;
  peer_tick_kernel<<<1, 1, 0, stream>>>(a->dev_state, a->beta1, a->beta2);
  TFPP_CHECK_LAUNCH();
  peer_barrier_kernel<<<1, 32, 0, stream>>>(p, 0);   // every rank's gradients are complete
  TFPP_CHECK_LAUNCH();
  if (hi4 > lo4) {
    const int grid = TFPP_NUM_SMS * 4;
#define LAUNCH(Wn)                                                                                                   \
  peer_adamw_kernel<Wn><<<grid, 256, 0, stream>>>(p, a->exp_avg, a->exp_avg_sq, a->max_exp_avg_sq, lo4, hi4, a->beta1, \
                                                  a->beta2, a->eps, a->weight_decay, a->dev_state, a->opt_flags)
    switch (a->world) {
      case 1: LAUNCH(1); break;
      case 2: LAUNCH(2); break;
      case 3: LAUNCH(3); break;
      case 4: LAUNCH(4); break;
      case 5: LAUNCH(5); break;
      case 6: LAUNCH(6); break;
      case 7: LAUNCH(7); break;
      default: LAUNCH(8); break;
    }
#undef LAUNCH
    TFPP_CHECK_LAUNCH();
  }
  peer_barrier_kernel<<<1, 32, 0, stream>>>(p, 1);   // every rank's parameter pushes have landed
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
