// The ONE exchange step of the training path (reference train.py:437-442: backward -> optimizer.step, data-parallel replicas
// average their gradients in between) as ONE kernel per rank over NVLink / NVSwitch peer memory:
//
//   gradient all-reduce  +  SGD-momentum update with the reference's per-layer groups  +  EMA of the new weights
//
// Every rank's flat fp32 gradient bucket lives in symmetric memory (the same allocation mapped into every peer, plus -- on an
// NVSwitch box -- one MULTICAST address for all of them).  A launch runs, on every rank concurrently:
//
//   B1   cross-GPU barrier: all ranks' backward passes have written their buckets (signal pads in peer memory, sequence numbers)
//   P1   in-switch all-reduce of the rank's own 1/W slice: `multimem.ld_reduce.add.v4.f32` returns the SUM over all ranks' buckets
//        (the NVSwitch adds, one load), `multimem.st.v4.f32` writes it back into ALL ranks' buckets (the switch replicates, one
//        store): reduce-scatter + all-gather without a byte crossing a link twice.  Without a multicast mapping the same slice is
//        summed from W peer pointers in rank order and stored to W peer pointers (plain P2P loads / stores).
//   B2   barrier: every slice of every rank has landed
//   P2   the whole optimizer step on the now reduced local bucket -- exactly sgd_ema_multi_kernel's arithmetic (train.cu), so
//        momentum buffers and EMA shadows stay replicated and bit-identical across ranks and checkpoints need no gather
//
// The launch is cooperative (as many 512-thread CTAs per SM as are co-resident, up to 3: CTAs spin at the barriers).  Each slice is reduced exactly
// once and then broadcast, so all ranks apply bit-identical gradients.  NCCL is not on this path (it still provides the rendezvous
// and the timing barrier of the host code).
#include <cooperative_groups.h>
#include "common.cuh"

namespace ppy {
namespace {

struct PeerStep {
  float* grad_local;                 // this rank's mapping of its own bucket
  float* grad_mc;                    // multicast address of the bucket (nullptr: use grad_peers)
  float* const* grad_peers;          // [world] device pointers: rank r's bucket as mapped here
  unsigned int* const* pads;         // [world] signal pads (uint32 words), rank r's pad as mapped here
  int pad_slot;                      // first word of the `world` words this kernel uses in every pad
  int rank, world;
  unsigned int seq;                  // barrier values 2*seq + 1 and 2*seq + 2 (monotonic over the run)
  unsigned int* done;                // [2]: local CTA counter (zero between launches), error word (set when a barrier wait timed out)
  long long total4;                  // bucket length in float4 (padded)
  float* const* params; float* mom; float* shadow; const long long* offsets; const long long* shadow_offsets;
  const float* lr_mult; const float* wd; int num_tensors;
  float lr, momentum, grad_scale; int first_step; float decay, one_minus_decay;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// thread 0 of every CTA waits until every rank has published `value` in this rank's pad.  A peer that never arrives (its process
// died, it skipped the call) must not hang the GPU: after kWaitTimeoutNs (3 s) the wait gives up and raises the error word behind the
// `done` counter (done[1]), which the host checks (Trainer.timing_summary / state_dict); the step's results are then invalid.
constexpr unsigned long long kWaitTimeoutNs = 3ull * 1000 * 1000 * 1000;
__device__ __forceinline__ void wait_all(const PeerStep& a, unsigned int value) {
  if (threadIdx.x == 0) {
    const unsigned int* mine = a.pads[a.rank] + a.pad_slot;
    const unsigned long long t0 = global_ns();
    for (int p = 0; p < a.world; ++p)
      while ((int)(ld_acquire_sys(mine + p) - value) < 0) {
        __nanosleep(64);
        if (global_ns() - t0 > kWaitTimeoutNs) { a.done[1] = 1u; break; }
      }
  }
  __syncthreads();
}
__device__ __forceinline__ void signal_all(const PeerStep& a, unsigned int value) {       // called by one warp
  for (int p = threadIdx.x & 31; p < a.world; p += 32) st_release_sys(a.pads[p] + a.pad_slot + a.rank, value);
}

__global__ void __launch_bounds__(512, 2) allreduce_sgd_ema_kernel(const PeerStep a) {
  // ---- B1: this rank's bucket is complete (stream order); tell everybody, wait for everybody
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    __threadfence_system();
    signal_all(a, 2u * a.seq + 1u);
  }
  wait_all(a, 2u * a.seq + 1u);
  // ---- P1: all-reduce of this rank's slice
  const long long per = (a.total4 + a.world - 1) / a.world;
  const long long lo = per * a.rank, hi = lo + per < a.total4 ? lo + per : a.total4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (a.grad_mc != nullptr) {
    // (four reductions in flight per thread were measured: no gain at 2 GPUs, 0.40 -> 0.45 ms at 8 -- the phase is bound by the
    // barriers and the rank skew, not by requests in flight)
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride)
      multimem_st(a.grad_mc + 4 * i, multimem_ld_reduce_add(a.grad_mc + 4 * i));
  } else {
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < a.world; ++p) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(a.grad_peers[p]) + i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      for (int p = 0; p < a.world; ++p) __stcg(reinterpret_cast<float4*>(a.grad_peers[p]) + i, s);
    }
  }
  // ---- B2: all of this rank's stores are out -> signal; wait until every rank's slice has landed here
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(a.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence_system();
    if (threadIdx.x == 0) *a.done = 0u;
    signal_all(a, 2u * a.seq + 2u);
  }
  wait_all(a, 2u * a.seq + 2u);
  // ---- P2: optimizer + EMA over every tensor, reading the reduced bucket (L2 reads: peers wrote it)
  for (int t = 0; t < a.num_tensors; ++t) {
    const long long begin = a.offsets[t], count = a.offsets[t + 1] - begin;
    float* __restrict__ p = a.params[t];
    const float* g = a.grad_local + begin;
    float* __restrict__ buf = a.mom + begin;
    float* __restrict__ sh = a.shadow ? a.shadow + a.shadow_offsets[t] : nullptr;
    const float lr_t = a.lr * a.lr_mult[t], wd_t = a.wd[t];
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < count; i0 += 4 * stride) {
      float w[4], gr[4], bu[4], so[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                      // four independent elements in flight per thread
        const long long i = i0 + u * stride;
        if (i < count) {
          w[u] = p[i]; gr[u] = __ldcg(g + i); bu[u] = a.first_step ? 0.f : buf[i]; so[u] = sh ? sh[i] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = i0 + u * stride;
        if (i >= count) continue;
        const float d = gr[u] * a.grad_scale + wd_t * w[u];
        const float b = a.first_step ? d : a.momentum * bu[u] + d;
        buf[i] = b;
        const float w2 = w[u] - lr_t * b;
        p[i] = w2;
        if (sh) sh[i] = __fadd_rn(__fmul_rn(a.decay, so[u]), __fmul_rn(a.one_minus_decay, w2));
      }
    }
  }
}

}  // namespace
}  // namespace ppy

extern "C" int ppy_allreduce_sgd_ema(float* grad_local, float* grad_multicast, float* const* grad_peers, unsigned int* const* signal_pads,
                                     int pad_slot, int rank, int world, unsigned int seq, unsigned int* done, long long total_padded,
                                     float* const* params, float* momentum_flat, float* shadow_flat, const long long* offsets,
                                     const long long* shadow_offsets, const float* lr_mult, const float* weight_decay, int num_tensors,
                                     float lr, float momentum, float grad_scale, int first_step, float ema_decay, float ema_one_minus_decay,
                                     ppy_stream_t s) {
  using namespace ppy;
  PPY_REQUIRE(grad_local && signal_pads && done && params && momentum_flat && offsets && lr_mult && weight_decay && num_tensors > 0);
  PPY_REQUIRE((grad_multicast || grad_peers) && world >= 1 && rank >= 0 && rank < world && pad_slot >= 0);
  PPY_REQUIRE(total_padded > 0 && total_padded % 4 == 0 && (reinterpret_cast<uintptr_t>(grad_local) & 15) == 0);
  PPY_REQUIRE((reinterpret_cast<uintptr_t>(grad_multicast) & 15) == 0 && (!shadow_flat || shadow_offsets));
  PeerStep a;
  a.grad_local = grad_local; a.grad_mc = grad_multicast; a.grad_peers = grad_peers; a.pads = signal_pads; a.pad_slot = pad_slot;
  a.rank = rank; a.world = world; a.seq = seq; a.done = done; a.total4 = total_padded / 4;
  a.params = params; a.mom = momentum_flat; a.shadow = shadow_flat; a.offsets = offsets; a.shadow_offsets = shadow_offsets;
  a.lr_mult = lr_mult; a.wd = weight_decay; a.num_tensors = num_tensors;
  a.lr = lr; a.momentum = momentum; a.grad_scale = grad_scale; a.first_step = first_step; a.decay = ema_decay; a.one_minus_decay = ema_one_minus_decay;
  int dev = 0, sms = 0, coop = 0;
  int rc = check_cuda(cudaGetDevice(&dev));
  if (rc) return rc;
  if ((rc = check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)))) return rc;
  if ((rc = check_cuda(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev)))) return rc;
  if (!coop) return PPY_ERR_UNSUPPORTED;
  int per_sm = 1;                                       // every CTA must be resident (they spin at the barriers): ask, do not assume
  if ((rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, allreduce_sgd_ema_kernel, 512, 0)))) return rc;
  if (per_sm < 1) return PPY_ERR_UNSUPPORTED;
  if (per_sm > 3) per_sm = 3;
  void* args[] = {(void*)&a};
  rc = check_cuda(cudaLaunchCooperativeKernel((const void*)allreduce_sgd_ema_kernel, dim3((unsigned)(sms * per_sm)), dim3(512), args, 0, as_stream(s)));
  if (rc) return rc;
  count_launch();
  return PPY_OK;
}
