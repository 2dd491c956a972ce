// K8 ("next" row N2 of SURVEY section 8f): cut the raw int16 signal of every read into the
// pieces that the chunk merging (K7) produced.
//
// Replaces the only computation inside split_f5.split_signal (catfish/split_f5.py:36,64):
//     new_signal = signal_dset[s[0] : s[1]]
// for a batch of (read, start, end) ranges, with numpy/h5py slice semantics (a negative bound
// counts from the end, bounds are clamped to the read, an empty or inverted range yields an empty
// piece).  Writing the pieces back into FAST5 containers (HDF5, gzip) stays on the host.
// Two launches: piece lengths (exclusive scan on the host side of the ABI is avoided by a
// single-block device scan), then one CTA per piece copies with 128-bit accesses where aligned.
#include "common.cuh"

namespace cf {

__device__ __forceinline__ void slice_bounds(int64_t s, int64_t e, int64_t len, int64_t* b0, int64_t* b1) {
    if (s < 0) s += len;
    if (e < 0) e += len;
    s = s < 0 ? 0 : (s > len ? len : s);
    e = e < 0 ? 0 : (e > len ? len : e);
    *b0 = s;
    *b1 = e > s ? e : s;
}

__global__ void k8_piece_lengths_kernel(const int64_t* __restrict__ ranges, const int32_t* __restrict__ range_read,
                                        const int64_t* __restrict__ offsets, int64_t n_ranges,
                                        int64_t* __restrict__ lengths) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ranges) return;
    const int r = range_read[i];
    int64_t b0, b1;
    slice_bounds(ranges[2 * i], ranges[2 * i + 1], offsets[r + 1] - offsets[r], &b0, &b1);
    lengths[i] = b1 - b0;
}

__global__ void __launch_bounds__(1024)
k8_scan_kernel(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    __shared__ long long warp_sums[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const long long v = i < n ? in[i] : 0;
        long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            long long o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        long long wbase = 0, tot = 0;
        for (int k = 0; k < 32; ++k) {
            long long s = warp_sums[k];
            if (k < warp) wbase += s;
            tot += s;
        }
        const long long c = carry;
        if (i < n) out[i] = c + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

__global__ void __launch_bounds__(256)
k8_copy_pieces_kernel(const int16_t* __restrict__ raw, const int64_t* __restrict__ ranges,
                      const int32_t* __restrict__ range_read, const int64_t* __restrict__ offsets,
                      const int64_t* __restrict__ piece_off, int64_t capacity, int16_t* __restrict__ out) {
    const int64_t i = blockIdx.x;
    const int r = range_read[i];
    int64_t b0, b1;
    slice_bounds(ranges[2 * i], ranges[2 * i + 1], offsets[r + 1] - offsets[r], &b0, &b1);
    const int16_t* src = raw + offsets[r] + b0;
    int64_t n = b1 - b0;
    const int64_t dst0 = piece_off[i];
    if (dst0 + n > capacity) n = capacity > dst0 ? capacity - dst0 : 0;
    int16_t* dst = out + dst0;
    // 128-bit body when source and destination share their alignment phase, 16-bit otherwise
    if (((reinterpret_cast<uintptr_t>(src) ^ reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        int64_t head = ((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15) >> 1;
        if (head > n) head = n;
        for (int64_t k = threadIdx.x; k < head; k += blockDim.x) dst[k] = src[k];
        const int64_t nv = (n - head) >> 3;
        const int4* s4 = reinterpret_cast<const int4*>(src + head);
        int4* d4 = reinterpret_cast<int4*>(dst + head);
        for (int64_t k = threadIdx.x; k < nv; k += blockDim.x) d4[k] = __ldg(s4 + k);
        for (int64_t k = head + (nv << 3) + threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
    } else {
        for (int64_t k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
    }
}

int k8_split_raw(const int16_t* raw, const int64_t* offsets_dev, const int64_t* ranges, const int32_t* range_read,
                 int64_t n_ranges, int64_t* lengths_scratch, int64_t* piece_off, int64_t capacity, int16_t* out,
                 cudaStream_t stream) {
    if (n_ranges <= 0) {
        CF_CUDA(cudaMemsetAsync(piece_off, 0, sizeof(int64_t), stream));
        return CF_OK;
    }
    k8_piece_lengths_kernel<<<(unsigned)ceil_div(n_ranges, 256), 256, 0, stream>>>(ranges, range_read, offsets_dev, n_ranges, lengths_scratch);
    CF_LAUNCHED();
    k8_scan_kernel<<<1, 1024, 0, stream>>>(lengths_scratch, n_ranges, piece_off);
    CF_LAUNCHED();
    k8_copy_pieces_kernel<<<(unsigned)n_ranges, 256, 0, stream>>>(raw, ranges, range_read, offsets_dev, piece_off, capacity, out);
    CF_LAUNCHED();
    return CF_OK;
}

}  // namespace cf
