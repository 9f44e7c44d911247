// tokens.cu — what the reference's consumers do with the semantic ids:
//   token = code + position * codebook_size + 1          (RQVAE-T5/data_read.ipynb cell 2: item_to_offset_code;
//                                                          token ranges per position: check_data_alignment.py:105-121)
//   sequence rows = tokens[item_id - 1] for 1-indexed item ids of the interaction lists (same cell).
// Integer, HBM-bound, bit-exact; int32 out because the reference stores the TIGER datasets as int32 (cell 3).
#include "common.cuh"

namespace rqb {
namespace {

__global__ void offset_tokens_kernel(const int64_t *__restrict__ ids, int64_t n, int C, int K, int32_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n * C; p += stride) {
        const int c = (int)(p % C);
        out[p] = (int32_t)(ids[p] + (int64_t)c * K + 1);
    }
}

// out[j, :] = tokens[item_ids[j] - 1, :]; ids outside [1, n] set *bad (numpy would raise IndexError / wrap silently for 0)
__global__ void gather_item_tokens_kernel(const int32_t *__restrict__ tokens, int64_t n, int C, const int64_t *__restrict__ item_ids,
                                          int64_t count, int32_t *__restrict__ out, int *__restrict__ bad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < count * C; p += stride) {
        const int64_t j = p / C;
        const int c = (int)(p - j * C);
        const int64_t it = item_ids[j];
        if (it < 1 || it > n) {
            if (c == 0) atomicExch(bad, 1);
            out[p] = 0;
        } else {
            out[p] = tokens[(it - 1) * C + c];
        }
    }
}

int grid_of(int64_t count) {
    int64_t b = (count + 255) / 256;
    if (b > kNumSMs * 8) b = kNumSMs * 8;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace
}  // namespace rqb

using namespace rqb;

extern "C" int rqb200_offset_tokens(const int64_t *ids_dev, int64_t n, int n_cols, int codebook_size, int32_t *tokens_dev,
                                    void *stream) {
    if (n == 0) return 0;
    RQB_CHECK(ids_dev && tokens_dev, "NULL buffer");
    RQB_CHECK(n_cols >= 1 && codebook_size >= 1, "bad shape");
    count_launch();
    offset_tokens_kernel<<<grid_of(n * n_cols), 256, 0, (cudaStream_t)stream>>>(ids_dev, n, n_cols, codebook_size, tokens_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rqb200_gather_item_tokens(const int32_t *tokens_dev, int64_t n, int n_cols, const int64_t *item_ids_dev,
                                         int64_t count, int32_t *out_dev, int *bad_flag_dev, void *stream) {
    if (count == 0) return 0;
    RQB_CHECK(tokens_dev && item_ids_dev && out_dev && bad_flag_dev, "NULL buffer");
    count_launch();
    gather_item_tokens_kernel<<<grid_of(count * n_cols), 256, 0, (cudaStream_t)stream>>>(tokens_dev, n, n_cols, item_ids_dev, count,
                                                                                       out_dev, bad_flag_dev);
    RQB_LAUNCH_CHECK();
    return 0;
}
