// Micro-benchmark: how fast can one producer thread per SM stream a row-major fp16 matrix (rows x 256) into shared
// memory with TMA, as a function of the box shape?  Mirrors the operand loads of fold_gemm_kernel:
//   mode 0: box 32 x 128 rows (64-byte rows, SWIZZLE_64B), k-blocks walked inside a row block  [what the kernels do]
//   mode 1: box 64 x 64 rows  (128-byte rows, SWIZZLE_128B), same bytes per op
//   mode 2: box 64 x 128 rows (128-byte rows, 16 KB per op)
//   mode 3: box 256 x 16 rows  (whole 512-byte rows, SWIZZLE_NONE), same bytes per op
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dct_autoencoder_b200/csrc tools/tma_bw.cu -o gpurun_out/tma_bw
#include <cstdio>
#include <cstdlib>
#include "tc_ptx.cuh"
namespace dcta { void set_error(const char*, ...) {} }
using namespace dcta;

constexpr int STAGES = 6;
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap map, int rows_per_op, int ops_per_row_block,
                                                       int k_step, int n_row_blocks, uint32_t op_bytes, int* sink) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full[STAGES];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // one thread both issues and consumes: keep STAGES ops in flight
        uint32_t issued = 0, done = 0;
        const int total_ops = ((n_row_blocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * ops_per_row_block;
        int rb = blockIdx.x, kb = 0;
        while (done < (uint32_t)total_ops) {
            while (issued < (uint32_t)total_ops && issued - done < STAGES) {
                const int s = issued % STAGES;
                mbar_expect_tx(&full[s], op_bytes);
                tma_load_3d(&map, &full[s], smem + s * 16384, kb * k_step, rb * rows_per_op, 0);
                ++issued;
                if (++kb == ops_per_row_block) { kb = 0; rb += gridDim.x; }
            }
            mbar_wait(&full[done % STAGES], (done / STAGES) & 1);
            ++done;
        }
        if (smem[threadIdx.x] == 123 && sink) *sink = 1;
    }
}

int main(int argc, char** argv) {
    const int64_t rows = 256ll * 3 * 4 * 256;      // the stacked rows of config 2 (786432), K = 256
    const int K = 256;
    __half* d;
    cudaMalloc(&d, rows * K * 2);
    cudaMemset(d, 0, rows * K * 2);
    auto enc = get_encode_fn();
    struct Mode { int box_k, box_rows; CUtensorMapSwizzle sw; const char* name; } modes[] = {
        {32, 128, CU_TENSOR_MAP_SWIZZLE_64B, "box 32k x 128 rows, 64 B rows, SW64 (8 KB/op)"},
        {64, 64, CU_TENSOR_MAP_SWIZZLE_128B, "box 64k x 64 rows, 128 B rows, SW128 (8 KB/op)"},
        {64, 128, CU_TENSOR_MAP_SWIZZLE_128B, "box 64k x 128 rows, 128 B rows, SW128 (16 KB/op)"},
        {256, 16, CU_TENSOR_MAP_SWIZZLE_NONE, "box 256k x 16 rows, 512 B rows, no swizzle (8 KB/op)"},
        {32, 112, CU_TENSOR_MAP_SWIZZLE_64B, "box 32k x 112 rows, 64 B rows, SW64 (7 KB/op)"},
    };
    for (auto& m : modes) {
        CUtensorMap map;
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 1};
        cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)rows * K * 2};
        cuuint32_t box[3] = {(cuuint32_t)m.box_k, (cuuint32_t)m.box_rows, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         m.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int ops_per_row_block = K / m.box_k;
        const int n_row_blocks = (int)(rows / m.box_rows);
        const uint32_t op_bytes = m.box_k * m.box_rows * 2;
        cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * 16384 + 1024);
        for (int grid : {148, 296}) {
            float best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0);
                stream_kernel<<<grid, 64, STAGES * 16384 + 1024>>>(map, m.box_rows, ops_per_row_block, m.box_k, n_row_blocks, op_bytes, nullptr);
                cudaEventRecord(e1);
                cudaError_t err = cudaEventSynchronize(e1);
                if (err != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(err)); return 1; }
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("%-58s grid %3d: %7.1f us  %6.0f GB/s\n", m.name, grid, best * 1e3, rows * K * 2 / (best * 1e-3) / 1e9);
        }
    }
    return 0;
}
