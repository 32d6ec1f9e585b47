// All-gather of per-pair fp32 scores over NVLink peer memory: the one collective of the
// doc-parallel scoring path, as our own kernel instead of an NCCL call.
// Every rank owns a receive buffer [world][count] and a flag array [world] in memory that its peers have mapped (CUDA IPC /
// torch symmetric memory); rank r writes its slice into slot r of EVERY peer's buffer with plain stores through the
// NVLink-mapped pointers (one CTA per destination), publishes the call's sequence number in slot r of that peer's flag
// array (release, system scope) and returns when all `world` flags of its OWN array carry the sequence number (acquire):
// the kernel that follows on the stream sees all world * count scores.  The exchange is a few KB: pure latency, one
// launch, no staging copy, no proxy thread.
#include "common.cuh"

namespace cair {

constexpr int P2P_MAXWORLD = 16;
struct P2pPtrs {
  float* recv[P2P_MAXWORLD];
  uint32_t* flags[P2P_MAXWORLD];
};

__global__ void __launch_bounds__(256) allgather_p2p_kernel(const float* __restrict__ send, int64_t count, P2pPtrs pp, int rank,
                                                            int world, uint32_t seq) {
  const int dst = blockIdx.x, tid = threadIdx.x;
  float* out = pp.recv[dst] + (int64_t)rank * count;
  if ((count & 3) == 0 && (((uintptr_t)out | (uintptr_t)send) & 15) == 0) {
    for (int64_t i = tid; i < count / 4; i += blockDim.x) reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(send)[i];
  } else {
    for (int64_t i = tid; i < count; i += blockDim.x) out[i] = send[i];
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pp.flags[dst] + rank), "r"(seq) : "memory");
  }
  if (dst == rank && tid < world) {
    const uint32_t* f = pp.flags[rank] + tid;
    uint32_t v, spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (++spins > (1u << 28)) __trap();   // a peer never arrived: fail the launch instead of hanging the GPU
    } while ((int32_t)(v - seq) < 0);
  }
}

int32_t allgather_scores_p2p(const float* send, int64_t count, const uint64_t* peer_recv, const uint64_t* peer_flags, int rank,
                             int world, uint32_t seq, cudaStream_t s) {
  if (world < 1 || world > P2P_MAXWORLD || rank < 0 || rank >= world) return fail(CAIR_ERR_BAD_ARG, "allgather_scores: bad rank / world");
  if (!send || !peer_recv || !peer_flags || count < 0) return fail(CAIR_ERR_BAD_ARG, "allgather_scores: null argument");
  P2pPtrs pp;
  for (int r = 0; r < P2P_MAXWORLD; ++r) {
    pp.recv[r] = r < world ? reinterpret_cast<float*>(peer_recv[r]) : nullptr;
    pp.flags[r] = r < world ? reinterpret_cast<uint32_t*>(peer_flags[r]) : nullptr;
  }
  CAIR_LAUNCH(allgather_p2p_kernel, (unsigned)world, 256, 0, s, send, count, pp, rank, world, seq);
  return CAIR_OK;
}

}  // namespace cair
