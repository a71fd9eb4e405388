// FT-ZeRO-1: the partitioned optimizer update fused into the cross-replica collectives (sm_100a).
//
// The reference all-reduces every gradient, divides by N on the host side and then runs the FULL
// optimizer on every replica, gated by a host-synchronous should_commit
// (/root/reference/torchft/manager.py:466-478, /root/reference/torchft/optim.py:52-55). Here the
// replicated dimension also partitions the optimizer:
//
//   backward :  zero1_reduce_scatter_kernel   per unit (layer), overlapped with backward.
//               Rank r reduces slice r of the unit straight out of every peer's HBM (P2P loads, or
//               ONE multimem.ld_reduce per 16 B when the segment is bound to an NVLS multicast
//               object), fp32 accumulate in fixed rank order, 1/num_participants scale and bf16
//               cast in registers, zero contribution of healing/spare replicas, and writes the
//               reduced slice to itself AND to its k-1 buddies (k-way replicated ownership, so a
//               replica that dies never holds the only copy of a shard of optimizer state).
//   commit   :  zero1_commit_kernel           ONE tiny kernel replaces the host sync + RPC: every
//               rank publishes "my step is clean" (no latched error, enough participants) into
//               its peers' signal pads, the verdict is the AND over the quorum, written to a
//               device gate word, a device step counter and a host-mapped ring the Manager reads
//               lazily at the next start_quorum.
//   update   :  zero1_adamw_allgather_kernel  per unit in forward order, gated on the device word.
//               AdamW (fp32 master/m/v) on the slices this rank holds; the PRIMARY holder packs
//               the new weights to bf16 and stores them into every replica's parameter buffer
//               (P2P stores, or ONE multimem.st per 16 B) -- the all-gather is the epilogue of the
//               update, tile by tile, and the next forward overlaps it unit by unit.
//
// Per rank and step this moves the same NVLink bytes as a two-shot all-reduce but divides the
// 28 B/param optimizer HBM traffic (and the state a replica must hold) by N/k.
#include <stdexcept>
#include <string>

#include "api.h"
#include "common.cuh"

namespace tft {

namespace {

using bf16 = __nv_bfloat16;
using P8 = Pack<bf16>;

__device__ __forceinline__ Vec16 mm_ld_reduce_bf16(const void* p) {
  Vec16 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void mm_st_bf16(void* p, const Vec16& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// Slice geometry shared by every kernel here and by parallel/zero1.py (must stay in sync):
// a unit of `nelem` bf16 elements (multiple of 8) = nvec 16-byte vectors, cut into W slices of
// ceil(nvec / W) vectors; slice s belongs to quorum rank s (primary) and ranks s+1..s+k-1 (buddies).
struct Geo {
  size_t nvec, slice;
  __device__ Geo(size_t nelem, int W) : nvec(nelem / 8), slice((nelem / 8 + W - 1) / W) {}
  __device__ size_t lo(int s) const { return min((size_t)s * slice, nvec); }
  __device__ size_t hi(int s) const { return min((size_t)(s + 1) * slice, nvec); }
};

struct RSArgs {
  PeerTable pt;       // data[] = every rank's gradient segment
  StatusBlock* st;
  char* mc;           // multicast VA of the gradient segment (nullptr = P2P loads)
  size_t off;         // byte offset of the unit inside the segment
  size_t nelem;       // elements in the unit (multiple of 8)
  float scale;
  uint64_t flag;      // consumes flag+1, flag+2
  int channel;
  int contribute;
  int replication;    // k
  int barrier_mode;
};

template <int W, bool NVLS>
__global__ void __launch_bounds__(512, 1) zero1_reduce_scatter_kernel(RSArgs a) {
  const int rank = a.pt.rank;
  const Geo g(a.nelem, W);
  bf16* mine = reinterpret_cast<bf16*>(reinterpret_cast<char*>(a.pt.data[rank]) + a.off);

  if (!a.contribute) {
    // healing / spare replica: its gradients must count as zeros (reference manager.py:441-442);
    // zero them in place instead of a separate zero_() pass. Block b zeroes chunk b of every slice.
    const size_t chunk = (g.slice + gridDim.x - 1) / gridDim.x;
    for (int s = 0; s < W; ++s) {
      const size_t lo = g.lo(s) + blockIdx.x * chunk, hi = min(lo + chunk, g.hi(s));
      for (size_t v = lo + threadIdx.x; v < hi; v += blockDim.x) st_stream(mine + v * 8, Vec16{0, 0, 0, 0});
    }
  }
  if (!block_barrier(a.pt, a.channel, a.flag + 1, a.st, /*release=*/!a.contribute, /*acquire=*/false, a.barrier_mode))
    return;

  {
    const bf16* src[W];
    bf16* dst[W];
#pragma unroll
    for (int p = 0; p < W; ++p) {
      src[p] = reinterpret_cast<const bf16*>(reinterpret_cast<const char*>(a.pt.data[p]) + a.off);
      dst[p] = reinterpret_cast<bf16*>(reinterpret_cast<char*>(a.pt.data[(rank + p) % W]) + a.off);
    }
    const char* mcbase = a.mc + a.off;
    const int k = min(a.replication, W);
    const size_t chunk = (g.slice + gridDim.x - 1) / gridDim.x;
    const size_t lo = g.lo(rank) + blockIdx.x * chunk, hi = min(lo + chunk, g.hi(rank));
    constexpr int U = NVLS ? 8 : ((W >= 8) ? 2 : (W >= 3 ? 4 : 8));
    for (size_t base = lo; base < hi; base += (size_t)blockDim.x * U) {
      if constexpr (NVLS) {
        Vec16 r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t v = base + threadIdx.x + (size_t)u * blockDim.x;
          if (v < hi) r[u] = mm_ld_reduce_bf16(mcbase + v * 16);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t v = base + threadIdx.x + (size_t)u * blockDim.x;
          if (v < hi) {
            float f[8];
            P8::unpack(r[u], f);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] *= a.scale;
            const Vec16 out = P8::pack(f);
#pragma unroll
            for (int j = 0; j < W; ++j)
              if (j < k) st_stream(dst[j] + v * 8, out);
          }
        }
      } else {
        Vec16 in[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t v = base + threadIdx.x + (size_t)u * blockDim.x;
          if (v < hi) {
#pragma unroll
            for (int p = 0; p < W; ++p) in[u][p] = ld_stream(src[p] + v * 8);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t v = base + threadIdx.x + (size_t)u * blockDim.x;
          if (v < hi) {
            float acc[8], f[8];
            P8::unpack(in[u][0], acc);
#pragma unroll
            for (int p = 1; p < W; ++p) {
              P8::unpack(in[u][p], f);
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[i] += f[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] *= a.scale;
            const Vec16 out = P8::pack(acc);
#pragma unroll
            for (int j = 0; j < W; ++j)
              if (j < k) st_stream(dst[j] + v * 8, out);
          }
        }
      }
    }
  }
  // buddies read the pushed slice in a LATER kernel (the gated update): release only
  block_barrier(a.pt, a.channel, a.flag + 2, a.st, /*release=*/true, /*acquire=*/false, a.barrier_mode);
}

// ---------------------------------------------------------------------------
// Commit verdict on the device (replaces current-stream synchronize + should_commit RPC for
// replica groups of one rank, /root/reference/torchft/manager.py:884-903).
// ---------------------------------------------------------------------------
struct CommitArgs {
  PeerTable pt;        // pads only
  StatusBlock* st;
  int* gate;           // device: gate[0] = verdict of this step, gate[1] = optimizer step count t
  uint64_t flag;       // consumes flag+1
  uint32_t seq;        // commit sequence number (slot in the host ring)
  int channel;
  int host_ok;         // host-side part of the verdict (enough participants, no host-latched error)
  int exchange;        // 1 = AND the verdict over the quorum through the signal pads
};

__global__ void __launch_bounds__(32, 1) zero1_commit_kernel(CommitArgs a) {
  const int t = threadIdx.x;
  const int W = a.pt.world, rank = a.pt.rank;
  const int local_ok = (a.host_ok && a.st->error == kOk && !a.st->abort) ? 1 : 0;
  int ok = local_ok;
  if (a.exchange && W > 1) {
    const uint64_t f = a.flag + 1;
    if (t < W && t != rank) {
      // slot 1 carries the verdict, slot 0 the arrival; a NOT-ok verdict is published too, so
      // clean peers learn immediately instead of waiting out their spin budget
      st_relaxed_sys(&a.pt.pads[t]->sig[a.channel][1][rank], f * 2 + (uint64_t)local_ok);
      fence_acq_rel_sys();
      st_release_sys(&a.pt.pads[t]->sig[a.channel][0][rank], f);
      const uint64_t* arrive = &a.pt.pads[rank]->sig[a.channel][0][t];
      bool got;
      if (a.st->error != kOk) {
        // our own step already failed: do not wait for anybody (their answer cannot change ours)
        got = false;
      } else {
        got = wait_flag(arrive, f, a.st, t, a.pt.timeout_ns);
      }
      if (got) {
        fence_acq_rel_sys();
        const uint64_t v = ld_relaxed_sys(&a.pt.pads[rank]->sig[a.channel][1][t]);
        ok = (v >= f * 2 && (v & 1ull)) ? local_ok : 0;  // stale (older) verdicts compare smaller
      } else {
        ok = 0;
      }
    }
    ok = __all_sync(0xffffffffu, ok);
  }
  if (t == 0) {
    a.gate[0] = ok;
    if (ok) a.gate[1] += 1;
    a.st->verdict[a.seq & (kVerdictRing - 1)] = (a.seq << 1) | (uint32_t)ok;
    __threadfence_system();
  }
}

// ---------------------------------------------------------------------------
// Gated AdamW on the held slices + all-gather of the new bf16 weights.
// ---------------------------------------------------------------------------
struct UpdArgs {
  PeerTable pt;        // data[] = every rank's PARAMETER segment
  StatusBlock* st;
  char* mc;            // multicast VA of the parameter segment (nullptr = P2P stores)
  const int* gate;     // gate[0] verdict, gate[1] step count (already incremented for this step)
  size_t poff;         // byte offset of the unit inside the parameter segment
  const bf16* grad;    // local reduced gradients of the unit
  float* master;       // local fp32 state of the unit (full-size arrays; only held slices are valid)
  float* m;
  float* v;
  size_t nelem;
  float lr, b1, b2, eps, wd;
  uint64_t flag;       // consumes flag+1
  int channel;
  int replication;
  int mode;            // 0 = gated update + push, 1 = refresh (push bf16(master) of primary slices, ungated)
  int barrier_mode;
};

template <int W, bool NVLS>
__global__ void __launch_bounds__(512, (W == 1) ? 2 : 1) zero1_adamw_allgather_kernel(UpdArgs a) {
  if (a.mode == 0 && a.gate[0] == 0) return;  // uniform over the grid AND (unanimous verdict) over the quorum
  const int rank = a.pt.rank;
  const Geo g(a.nelem, W);
  const int k = min(a.replication, W);
  const float tf = (float)a.gate[1];
  const float bc1 = 1.f - powf(a.b1, tf), bc2 = 1.f - powf(a.b2, tf);
  const float step_size = a.lr / bc1;
  const float inv_bc2_sqrt = rsqrtf(bc2);
  const float decay = 1.f - a.lr * a.wd;
  bf16* dst[W];
#pragma unroll
  for (int p = 0; p < W; ++p)
    dst[p] = reinterpret_cast<bf16*>(reinterpret_cast<char*>(a.pt.data[(rank + p) % W]) + a.poff);
  char* mcbase = a.mc + a.poff;
  const size_t chunk = (g.slice + gridDim.x - 1) / gridDim.x;

  for (int j = 0; j < (a.mode == 0 ? k : 1); ++j) {
    const int s = (rank - j + W) % W;  // j = 0: primary slice, j > 0: slices this rank backs up
    const size_t lo = g.lo(s) + blockIdx.x * chunk, hi = min(lo + chunk, g.hi(s));
    for (size_t v = lo + threadIdx.x; v < hi; v += blockDim.x) {
      float w[8];
      const Vec16 w0 = ld_stream(a.master + v * 8), w1 = ld_stream(a.master + v * 8 + 4);
      if (a.mode == 0) {
        // all 7 x 16 B loads in flight before any use
        const Vec16 gv = ld_stream(a.grad + v * 8);
        const Vec16 m0 = ld_stream(a.m + v * 8), m1 = ld_stream(a.m + v * 8 + 4);
        const Vec16 v0 = ld_stream(a.v + v * 8), v1 = ld_stream(a.v + v * 8 + 4);
        float gr[8], m[8], vv[8];
        P8::unpack(gv, gr);
        Pack<float>::unpack(w0, w);
        Pack<float>::unpack(w1, w + 4);
        Pack<float>::unpack(m0, m);
        Pack<float>::unpack(m1, m + 4);
        Pack<float>::unpack(v0, vv);
        Pack<float>::unpack(v1, vv + 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          m[i] = a.b1 * m[i] + (1.f - a.b1) * gr[i];
          vv[i] = a.b2 * vv[i] + (1.f - a.b2) * gr[i] * gr[i];
          const float denom = sqrtf(vv[i]) * inv_bc2_sqrt + a.eps;
          w[i] = w[i] * decay - step_size * __fdividef(m[i], denom);
        }
        st_stream(a.master + v * 8, Pack<float>::pack(w));
        st_stream(a.master + v * 8 + 4, Pack<float>::pack(w + 4));
        st_stream(a.m + v * 8, Pack<float>::pack(m));
        st_stream(a.m + v * 8 + 4, Pack<float>::pack(m + 4));
        st_stream(a.v + v * 8, Pack<float>::pack(vv));
        st_stream(a.v + v * 8 + 4, Pack<float>::pack(vv + 4));
      } else {
        Pack<float>::unpack(w0, w);
        Pack<float>::unpack(w1, w + 4);
      }
      if (j == 0) {
        const Vec16 out = P8::pack(w);
        if constexpr (NVLS) {
          mm_st_bf16(mcbase + v * 16, out);
        } else {
#pragma unroll
          for (int p = 0; p < W; ++p) st_stream(dst[p] + v * 8, out);
        }
      }
    }
  }
  if (W > 1)
    block_barrier(a.pt, a.channel, a.flag + 1, a.st, /*release=*/true, /*acquire=*/false, a.barrier_mode);
}

template <bool NVLS>
void launch_rs(const RSArgs& a, int blocks, int threads, cudaStream_t s) {
  switch (a.pt.world) {
    case 2: zero1_reduce_scatter_kernel<2, NVLS><<<blocks, threads, 0, s>>>(a); break;
    case 3: zero1_reduce_scatter_kernel<3, NVLS><<<blocks, threads, 0, s>>>(a); break;
    case 4: zero1_reduce_scatter_kernel<4, NVLS><<<blocks, threads, 0, s>>>(a); break;
    case 5: zero1_reduce_scatter_kernel<5, NVLS><<<blocks, threads, 0, s>>>(a); break;
    case 6: zero1_reduce_scatter_kernel<6, NVLS><<<blocks, threads, 0, s>>>(a); break;
    case 7: zero1_reduce_scatter_kernel<7, NVLS><<<blocks, threads, 0, s>>>(a); break;
    case 8: zero1_reduce_scatter_kernel<8, NVLS><<<blocks, threads, 0, s>>>(a); break;
    default: throw std::runtime_error("zero1_reduce_scatter: world size must be in [2, 8]");
  }
}

template <bool NVLS>
void launch_upd(const UpdArgs& a, int blocks, int threads, cudaStream_t s) {
  switch (a.pt.world) {
    case 1: zero1_adamw_allgather_kernel<1, false><<<blocks, threads, 0, s>>>(a); break;
    case 2: zero1_adamw_allgather_kernel<2, NVLS><<<blocks, threads, 0, s>>>(a); break;
    case 3: zero1_adamw_allgather_kernel<3, NVLS><<<blocks, threads, 0, s>>>(a); break;
    case 4: zero1_adamw_allgather_kernel<4, NVLS><<<blocks, threads, 0, s>>>(a); break;
    case 5: zero1_adamw_allgather_kernel<5, NVLS><<<blocks, threads, 0, s>>>(a); break;
    case 6: zero1_adamw_allgather_kernel<6, NVLS><<<blocks, threads, 0, s>>>(a); break;
    case 7: zero1_adamw_allgather_kernel<7, NVLS><<<blocks, threads, 0, s>>>(a); break;
    case 8: zero1_adamw_allgather_kernel<8, NVLS><<<blocks, threads, 0, s>>>(a); break;
    default: throw std::runtime_error("zero1_adamw_allgather: world size must be in [1, 8]");
  }
}

}  // namespace

void zero1_reduce_scatter_launch(const PeerTable& pt, StatusBlock* st, void* mc_base, size_t off, size_t nelem,
                                 float scale, uint64_t flag, int channel, int contribute, int replication,
                                 int blocks, int threads, int barrier_mode, cudaStream_t stream) {
  if (blocks < 1 || blocks > kMaxBlocks) throw std::runtime_error("zero1_reduce_scatter: bad grid");
  if (threads < 32 || threads > 512 || (threads & 31)) throw std::runtime_error("zero1_reduce_scatter: bad block size");
  if ((off & 15) || (nelem & 7)) throw std::runtime_error("zero1_reduce_scatter: unit must be 16 B aligned and a multiple of 8 elements");
  if (replication < 1) throw std::runtime_error("zero1_reduce_scatter: replication must be >= 1");
  RSArgs a{pt, st, reinterpret_cast<char*>(mc_base), off, nelem, scale, flag, channel, contribute, replication, barrier_mode};
  if (mc_base != nullptr)
    launch_rs<true>(a, blocks, threads, stream);
  else
    launch_rs<false>(a, blocks, threads, stream);
  TFT_CUDA_CHECK(cudaGetLastError());
}

void zero1_commit_launch(const PeerTable& pt, StatusBlock* st, int* gate, uint64_t flag, uint32_t seq, int channel,
                         int host_ok, int exchange, cudaStream_t stream) {
  CommitArgs a{pt, st, gate, flag, seq, channel, host_ok, exchange};
  zero1_commit_kernel<<<1, 32, 0, stream>>>(a);
  TFT_CUDA_CHECK(cudaGetLastError());
}

void zero1_adamw_allgather_launch(const PeerTable& pt, StatusBlock* st, void* mc_base, const int* gate, size_t poff,
                                  const void* grad, float* master, float* m, float* v, size_t nelem, float lr,
                                  float b1, float b2, float eps, float wd, uint64_t flag, int channel,
                                  int replication, int mode, int blocks, int threads, int barrier_mode,
                                  cudaStream_t stream) {
  if (blocks < 1 || (pt.world > 1 && blocks > kMaxBlocks)) throw std::runtime_error("zero1_adamw_allgather: bad grid");
  if (threads < 32 || threads > 512 || (threads & 31)) throw std::runtime_error("zero1_adamw_allgather: bad block size");
  if ((poff & 15) || (nelem & 7)) throw std::runtime_error("zero1_adamw_allgather: unit must be 16 B aligned and a multiple of 8 elements");
  UpdArgs a{pt, st, reinterpret_cast<char*>(mc_base), gate, poff, reinterpret_cast<const bf16*>(grad), master, m, v,
            nelem, lr, b1, b2, eps, wd, flag, channel, replication, mode, barrier_mode};
  if (mc_base != nullptr)
    launch_upd<true>(a, blocks, threads, stream);
  else
    launch_upd<false>(a, blocks, threads, stream);
  TFT_CUDA_CHECK(cudaGetLastError());
}

}  // namespace tft
