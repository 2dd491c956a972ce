// K7 ("next" row N1 of SURVEY section 8f): merge a read's homopolymer intervals into chunks of at
// least chunk_size samples, centre-pad short chunks, and list the non-homopolymer complement.
//
// Replaces the body of the per-file loop in the reference's CLI (catfish/catfish:58-81) and
// center_hp (catfish/catfish:121-135).  The reference works on Python lists that ALIAS the
// elements of hp_positions (merged_positions holds references, and `hp_positions[i - 1]` wraps to
// the last interval when i == 0); the kernel reproduces that literally: the read's intervals are
// mutated in place and the merged list is a list of indices into them.  center_hp's overflow
// branch (`[0] -= len_read - [1]`, which moves the start to the right) is reproduced as written.
// One thread per read (the scan over a read's intervals is sequential by nature, reads are
// independent); integer-only, bit-exact.
#include "common.cuh"

namespace cf {

__device__ __forceinline__ void center_hp(int64_t* m, int64_t len_read, int64_t chunk) {
    const int64_t len_hp = m[1] - m[0];
    if (len_hp < chunk) {
        const int64_t pad = chunk - len_hp;
        const int64_t left = pad >= 0 ? pad / 2 : -((-pad + 1) / 2);      // Python floor division
        const int64_t right = pad - left;
        m[0] -= left;
        m[1] += right;
        if (m[0] < 0) { m[1] -= m[0]; m[0] = 0; }
        if (m[1] > len_read) { m[0] -= len_read - m[1]; m[1] = len_read; }
    }
}

__global__ void k7_merge_chunks_kernel(int64_t* __restrict__ work, const int64_t* __restrict__ ioff,
                                       const int64_t* __restrict__ read_len, int n_reads, int64_t chunk,
                                       int32_t* __restrict__ idx, int64_t* __restrict__ merged,
                                       int64_t* __restrict__ merged_cnt, int64_t* __restrict__ nonhp,
                                       int64_t* __restrict__ nonhp_cnt) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const int64_t beg = ioff[r], n = ioff[r + 1] - beg;
    const int64_t len = read_len[r];
    int64_t* H = work + 2 * beg;                      // this read's intervals, mutated in place
    int32_t* mi = idx + beg + r;                      // merged list: indices into H, at most n + 1
    int64_t* M = merged + 2 * (beg + r);
    int64_t* N = nonhp + 2 * (beg + 2 * r);
    if (n == 0) {                                      // catfish:80-81: the whole read is non-HP
        merged_cnt[r] = 0;
        nonhp_cnt[r] = -1;                             // marks the reference's odd [([(0, len), len])] entry
        N[0] = 0;
        N[1] = len;
        return;
    }
    int64_t nm = 0;
    mi[nm++] = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t* last = H + 2 * mi[nm - 1];
        if (H[2 * i + 1] >= chunk + last[0]) {
            const int64_t prev = i == 0 ? n - 1 : i - 1;           // Python's hp_positions[-1]
            last[1] = H[2 * prev + 1];
            center_hp(last, len, chunk);
            mi[nm++] = (int32_t)i;
        }
    }
    center_hp(H + 2 * mi[nm - 1], len, chunk);
    for (int64_t k = 0; k < nm; ++k) { M[2 * k] = H[2 * mi[k]]; M[2 * k + 1] = H[2 * mi[k] + 1]; }
    merged_cnt[r] = nm;
    int64_t nn = 0, m_start = 0;
    for (int64_t k = 0; k < nm; ++k) {
        if (M[2 * k] > m_start) { N[2 * nn] = m_start; N[2 * nn + 1] = M[2 * k] - 1; ++nn; }
        m_start = M[2 * k + 1];
    }
    if (M[2 * (nm - 1) + 1] != len) { N[2 * nn] = M[2 * (nm - 1) + 1]; N[2 * nn + 1] = len; ++nn; }
    nonhp_cnt[r] = nn;
}

int k7_merge_chunks(int64_t* work, const int64_t* ioff, const int64_t* read_len, int n_reads, int64_t chunk,
                    int32_t* idx, int64_t* merged, int64_t* merged_cnt, int64_t* nonhp, int64_t* nonhp_cnt,
                    cudaStream_t stream) {
    if (n_reads <= 0) return CF_OK;
    k7_merge_chunks_kernel<<<(unsigned)ceil_div(n_reads, 128), 128, 0, stream>>>(work, ioff, read_len, n_reads, chunk, idx,
                                                                                 merged, merged_cnt, nonhp, nonhp_cnt);
    CF_LAUNCHED();
    return CF_OK;
}

}  // namespace cf
