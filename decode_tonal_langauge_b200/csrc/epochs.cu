// K8 onset-indexed epoch gather (bit-exact copy).
//
// out[n, c, 0:L] = src[c, start[n] : start[n]+L].  Works on 32-bit words so float32,
// float64 and integer sources all take the same path (elem_bytes 4 or 8).  The source
// window is misaligned by start[n] mod 4 words, the destination row is not: each
// thread gathers four consecutive words through L1 and issues one 128-bit store when
// the destination allows it.  HBM-bound: 8 B per gathered element.
#include "common.cuh"

namespace ecog {

constexpr int kGatherWarps = 8;

// one warp per (event, channel) row; rows are distributed grid-stride
__global__ void __launch_bounds__(kGatherWarps * 32)
epoch_gather_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ out, int64_t C,
                    int64_t ld_words, const int64_t* __restrict__ start, int64_t N, int64_t Lw,
                    int words_per_elem, bool vec_store) {
    const int lane = threadIdx.x & 31;
    const int64_t rows = N * C;
    for (int64_t r = (int64_t)blockIdx.x * kGatherWarps + (threadIdx.x >> 5); r < rows;
         r += (int64_t)gridDim.x * kGatherWarps) {
        const int64_t n = r / C, c = r - n * C;
        const uint32_t* s = src + c * ld_words + start[n] * words_per_elem;
        uint32_t* d = out + r * Lw;
        if (vec_store) {
            for (int64_t i = 4 * lane; i < Lw; i += 128) {   // Lw % 4 == 0 here
                uint4 v;
                v.x = s[i]; v.y = s[i + 1]; v.z = s[i + 2]; v.w = s[i + 3];
                *reinterpret_cast<uint4*>(d + i) = v;
            }
        } else {
            for (int64_t i = lane; i < Lw; i += 32) d[i] = s[i];
        }
    }
}

// K8b channel selection of an epoch tensor: out[n, j, :] = src[n, channels[j], :] (bit copy).
// One warp per (event, selected channel) row, same word-wise copy as the gather.
__global__ void __launch_bounds__(kGatherWarps * 32)
channel_select_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ out, int64_t N, int64_t C,
                      int64_t K, int64_t Lw, const int32_t* __restrict__ channels, bool vec) {
    const int lane = threadIdx.x & 31;
    const int64_t rows = N * K;
    for (int64_t r = (int64_t)blockIdx.x * kGatherWarps + (threadIdx.x >> 5); r < rows;
         r += (int64_t)gridDim.x * kGatherWarps) {
        const int64_t n = r / K, j = r - n * K;
        const uint32_t* s = src + (n * C + channels[j]) * Lw;
        uint32_t* d = out + r * Lw;
        if (vec) {
            for (int64_t i = 4 * lane; i < Lw; i += 128)
                *reinterpret_cast<uint4*>(d + i) = *reinterpret_cast<const uint4*>(s + i);
        } else {
            for (int64_t i = lane; i < Lw; i += 32) d[i] = s[i];
        }
    }
}

// 1- and 2-byte element types (int16 audio / raw ADC counts): the same copies, byte granular.
// row_bytes = L * elem_bytes; src row r starts at src_off(r) bytes.  Not a bandwidth path.
__global__ void __launch_bounds__(kGatherWarps * 32)
gather_bytes_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ out, int64_t rows, int64_t inner,
                    int64_t src_outer_stride, int64_t src_inner_stride, const int64_t* __restrict__ start,
                    const int32_t* __restrict__ channels, int64_t elem_bytes, int64_t row_bytes) {
    // epoch gather : rows = N*C, inner = C,  row (n, c) <- src + c*src_inner_stride + start[n]*elem_bytes
    // channel pick : rows = N*K, inner = K,  row (n, j) <- src + n*src_outer_stride + channels[j]*src_inner_stride
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * kGatherWarps + (threadIdx.x >> 5); r < rows;
         r += (int64_t)gridDim.x * kGatherWarps) {
        const int64_t n = r / inner, c = r - n * inner;
        const uint8_t* s = start ? src + c * src_inner_stride + start[n] * elem_bytes
                                 : src + n * src_outer_stride + (int64_t)channels[c] * src_inner_stride;
        uint8_t* d = out + r * row_bytes;
        for (int64_t i = lane; i < row_bytes; i += 32) d[i] = s[i];
    }
}

}  // namespace ecog

using namespace ecog;

extern "C" int ecog_channel_select(const void* d_src, void* d_out, int64_t N, int64_t C, int64_t L,
                                   const int32_t* d_channels, const int32_t* h_channels, int64_t K,
                                   int32_t elem_bytes, ecog_stream_t stream) {
    if (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4 && elem_bytes != 8)
        return fail(ECOG_E_VALUE, "ecog_channel_select: elem_bytes must be 1, 2, 4 or 8");
    if (N < 0 || C <= 0 || L <= 0 || K < 0) return fail(ECOG_E_VALUE, "ecog_channel_select: bad shape");
    for (int64_t j = 0; j < K; ++j)
        if (h_channels[j] < 0 || h_channels[j] >= C)
            return fail(ECOG_E_VALUE, "index %d is out of bounds for axis 1 with size %lld", h_channels[j], (long long)C);
    if (N == 0 || K == 0) return ECOG_OK;
    if (elem_bytes < 4) {
        int64_t nb = ceil_div(N * K, kGatherWarps);
        if (nb > (int64_t)kNumSMs * 32) nb = (int64_t)kNumSMs * 32;
        gather_bytes_kernel<<<(unsigned)nb, kGatherWarps * 32, 0, (cudaStream_t)stream>>>(
            (const uint8_t*)d_src, (uint8_t*)d_out, N * K, K, C * L * elem_bytes, L * elem_bytes, nullptr, d_channels,
            elem_bytes, L * elem_bytes);
        return check_launch("channel_select_bytes");
    }
    const int64_t Lw = L * (elem_bytes / 4);
    const bool vec = aligned16(d_src) && aligned16(d_out) && (Lw % 4 == 0);
    int64_t blocks = ceil_div(N * K, kGatherWarps);
    if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
    channel_select_kernel<<<(unsigned)blocks, kGatherWarps * 32, 0, (cudaStream_t)stream>>>(
        (const uint32_t*)d_src, (uint32_t*)d_out, N, C, K, Lw, d_channels, vec);
    return check_launch("channel_select");
}


extern "C" int ecog_epoch_gather(const void* d_src, void* d_out, int64_t C, int64_t T, int64_t ld,
                                 const int64_t* d_start, const int64_t* h_start, int64_t N, int64_t L,
                                 int32_t elem_bytes, ecog_stream_t stream) {
    if (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4 && elem_bytes != 8)
        return fail(ECOG_E_VALUE, "ecog_epoch_gather: elem_bytes must be 1, 2, 4 or 8");
    if (C <= 0 || T <= 0 || ld < T || L <= 0 || N < 0)
        return fail(ECOG_E_VALUE, "ecog_epoch_gather: bad shape");
    for (int64_t n = 0; n < N; ++n) {
        if (h_start[n] < 0)
            return fail(ECOG_E_VALUE, "Epoch %lld starts before the recording (start index %lld).",
                        (long long)n, (long long)h_start[n]);
        if (h_start[n] + L > T)
            return fail(ECOG_E_VALUE,
                        "Requested sample length exceeds data length. Start: %lld, End: %lld; Data length: %lld.",
                        (long long)h_start[n], (long long)(h_start[n] + L), (long long)T);
    }
    if (N == 0) return ECOG_OK;
    if (elem_bytes < 4) {
        int64_t nb = ceil_div(N * C, kGatherWarps);
        if (nb > (int64_t)kNumSMs * 32) nb = (int64_t)kNumSMs * 32;
        gather_bytes_kernel<<<(unsigned)nb, kGatherWarps * 32, 0, (cudaStream_t)stream>>>(
            (const uint8_t*)d_src, (uint8_t*)d_out, N * C, C, 0, ld * elem_bytes, d_start, nullptr, elem_bytes,
            L * elem_bytes);
        return check_launch("epoch_gather_bytes");
    }
    const int wpe = elem_bytes / 4;
    const int64_t Lw = L * wpe;
    const bool vec_store = aligned16(d_out) && (Lw % 4 == 0);
    int64_t blocks = ceil_div(N * C, kGatherWarps);
    if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
    epoch_gather_kernel<<<(unsigned)blocks, kGatherWarps * 32, 0, (cudaStream_t)stream>>>(
        (const uint32_t*)d_src, (uint32_t*)d_out, C, ld * wpe, d_start, N, Lw, wpe, vec_store);
    ECOG_TRY(check_launch("epoch_gather"));
    return ECOG_OK;
}
