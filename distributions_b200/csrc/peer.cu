// peer.cu -- device-side signalling between the ranks of a feature-sharded kind (one process per GPU, buffers
// exchanged as CUDA IPC handles).  Replaces host-blocking barriers around the NVLink push: after its score +
// push kernel a rank raises an epoch flag in every owner's memory (system-scope release), and an owner's
// sampler is preceded by a wait on the flags of all pushers (system-scope acquire).  Both are stream-ordered
// single-block kernels: a step never returns to the host between push and sample.
#include "common.cuh"

namespace distb200 {

struct PeerFlags {
    int n;
    uint32_t *ptr[kMaxPushOwners];
};

// thread r raises flag[index] = epoch in peer r's flag array.  The preceding kernel on this stream has completed,
// so its peer stores are performed with respect to this GPU; the system-scope fence orders them before the flag
// for any observer that acquires it.
__global__ void peer_signal_kernel(const PeerFlags f, int index, uint32_t epoch) {
    const int r = threadIdx.x;
    if (r >= f.n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(f.ptr[r] + index), "r"(epoch) : "memory");
}

// thread r waits until flags[r] >= epoch (epochs only grow); traps instead of hanging the device on a lost peer
__global__ void peer_wait_kernel(const uint32_t *flags, int n, uint32_t epoch) {
    const int r = threadIdx.x;
    if (r >= n) return;
    uint32_t v = 0;
    unsigned long long spins = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(flags + r) : "memory");
        if (static_cast<int32_t>(v - epoch) >= 0) break;
        if (++spins > (1ull << 26)) __trap();  // > 13 s of 200 ns naps
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace distb200

using namespace distb200;

extern "C" int dist_b200_peer_signal(dist_b200_ctx *ctx, void *const *flag_ptrs, int n_peers, int my_index, uint32_t epoch,
                                     void *stream) {
    if (!ctx || !flag_ptrs || n_peers < 1 || my_index < 0) return DIST_B200_ERR_INVALID;
    if (n_peers > kMaxPushOwners) return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "peer_signal: more than 16 peers");
    PeerFlags f{};
    f.n = n_peers;
    for (int i = 0; i < n_peers; ++i) {
        if (!flag_ptrs[i]) return fail(ctx, DIST_B200_ERR_INVALID, "peer_signal: null flag pointer");
        f.ptr[i] = static_cast<uint32_t *>(flag_ptrs[i]);
    }
    peer_signal_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(f, my_index, epoch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("peer_signal launch: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}

extern "C" int dist_b200_peer_wait(dist_b200_ctx *ctx, const void *flags_dev, int n_peers, uint32_t epoch, void *stream) {
    if (!ctx || !flags_dev || n_peers < 1) return DIST_B200_ERR_INVALID;
    if (n_peers > 32) return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "peer_wait: more than 32 peers");
    peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint32_t *>(flags_dev), n_peers, epoch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("peer_wait launch: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}
