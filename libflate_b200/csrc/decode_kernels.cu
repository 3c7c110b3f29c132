// decode_kernels.cu -- DEFLATE decode hot path (sm_100a).
//   k_inflate_streams : one warp per stream, blocks in order.  Every lane runs the scalar decode of
//                       inflate_core.cuh uniformly (broadcast loads, decode tables in shared memory);
//                       LZ77 copies and stored-block copies are spread over the 32 lanes.
// Reference behaviour: src/deflate/decode.rs:81-165, src/deflate/symbol.rs:193-243,387-484,
// src/huffman.rs:96-179, libflate_lz77/src/lib.rs:164-194.
#include "common.cuh"
#include "inflate_core.cuh"
#include "decode_dev.cuh"

namespace b2f {

struct WarpSync { __device__ __forceinline__ void operator()() const { __syncwarp(); } };

// Output policy: literals by lane 0, copies striped over the warp.  __syncwarp() orders the warp's earlier
// global stores before the reads of a copy (CUDA guarantees memory ordering among the participating lanes).
struct WarpOut {
    uint8_t *o; uint64_t capacity; uint32_t lane;
    __device__ __forceinline__ uint64_t cap() const { return capacity; }
    __device__ __forceinline__ void lit(uint64_t pos, uint8_t b) { if (lane == 0) o[pos] = b; }
    __device__ __forceinline__ void copy(uint64_t pos, uint32_t len, uint32_t dist) {
        __syncwarp();
        const uint8_t *src = o + pos - dist;
        if (dist >= len) { for (uint32_t k = lane; k < len; k += 32) o[pos + k] = src[k]; }
        else { for (uint32_t k = lane; k < len; k += 32) o[pos + k] = src[k % dist]; }
    }
    __device__ __forceinline__ void raw(uint64_t pos, const uint8_t *src, uint64_t n) {
        for (uint64_t k = lane; k < n; k += 32) o[pos + k] = src[k];
    }
};

constexpr uint32_t kInfWarps = 4;

__global__ void __launch_bounds__(kInfWarps * 32) k_inflate_streams(DecDev D) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    InflateTables *tabs = reinterpret_cast<InflateTables *>(smem_raw);
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t s = blockIdx.x * kInfWarps + wid;
    if (s >= D.n) return;
    InflateTables &T = tabs[wid];
    BitIn b;
    bi_init(b, D.in + D.in_off[s], D.in_len[s], 0);
    WarpOut out = { D.out + D.out_off[s], D.out_cap[s], lane };
    InflateResult R;
    inflate_blocks(b, T, out, 0, 0, 0xFFFFFFFFu, (int)lane, 32, WarpSync(), R);
    if (lane == 0) { D.status[s] = R.status; D.out_len[s] = R.out_len; D.consumed[s] = R.consumed; }
}

cudaError_t dec_init_attributes() {
    return cudaFuncSetAttribute(k_inflate_streams, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kInfWarps * sizeof(InflateTables)));
}
cudaError_t dec_launch_serial(const DecDev &D, cudaStream_t st) {
    if (D.n == 0) return cudaSuccess;
    k_inflate_streams<<<(D.n + kInfWarps - 1) / kInfWarps, kInfWarps * 32, kInfWarps * sizeof(InflateTables), st>>>(D);
    return cudaGetLastError();
}

}  // namespace b2f
