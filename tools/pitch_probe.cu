// Does the DRAM system care that every 16-bit operand box of this library is 32 rows x 128 bytes at a
// 1.5-6 KB row pitch (row-major activations), instead of one contiguous 4 KB burst (a K-blocked layout)?
// Copies the same number of bytes box by box, in the order the GEMMs walk them (a 128-row block, then
// its 64-column boxes left to right), for several row pitches; pitch 128 = the boxes are contiguous.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pitch_probe tools/pitch_probe.cu && ./pitch_probe
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

// One warp copies one box per iteration: 32 rows x 128 B = 8 warp-wide 16-byte accesses (4 rows each).
// blocked == 1: box (rb, cb) lives at ((rb * col_blocks + cb) * 4096) bytes, rows 128 B apart.
// blocked == 0: row-major, row pitch = col_blocks * 128 B.
__global__ void __launch_bounds__(256) copy_boxes(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t row_blocks,
                                                  int col_blocks, int blocked, int do_write) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t n_boxes = row_blocks * col_blocks;
  uint4 acc = make_uint4(0, 0, 0, 0);
  // a warp owns whole 32-row blocks and walks their column boxes in order (the k loop of a GEMM tile)
  for (int64_t b = warp * col_blocks; b < n_boxes; b += n_warps * col_blocks) {
    const int64_t rb = b / col_blocks;
    for (int cb = 0; cb < col_blocks; ++cb) {
      uint4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int row = 4 * k + (lane >> 3), piece = lane & 7;
        const int64_t idx = blocked ? ((rb * col_blocks + cb) * 256 + row * 8 + piece)
                                    : ((rb * 32 + row) * (int64_t)col_blocks * 8 + cb * 8 + piece);
        v[k] = src[idx];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int row = 4 * k + (lane >> 3), piece = lane & 7;
        const int64_t idx = blocked ? ((rb * col_blocks + cb) * 256 + row * 8 + piece)
                                    : ((rb * 32 + row) * (int64_t)col_blocks * 8 + cb * 8 + piece);
        if (do_write) dst[idx] = v[k];
        else { acc.x ^= v[k].x; acc.y ^= v[k].y; acc.z ^= v[k].z; acc.w ^= v[k].w; }
      }
    }
  }
  if (!do_write && (acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) dst[0] = acc;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  const int64_t bytes = (int64_t)6 << 30;                    // 6 GiB per buffer: far beyond the 126 MB L2
  uint4 *src, *dst;
  if (cudaMalloc(&src, bytes) != cudaSuccess || cudaMalloc(&dst, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(src, 1, bytes);
  cudaMemset(dst, 0, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int col_blocks_list[] = {12, 36, 48};               // H = 768 (1.5 KB pitch), 3H (4.5 KB), I (6 KB)
  for (int do_write = 0; do_write <= 1; ++do_write)
    for (int cbs : col_blocks_list)
      for (int blocked = 0; blocked <= 1; ++blocked)
        for (int occ = 2; occ <= 8; occ *= 2) {
          const int64_t row_blocks = bytes / ((int64_t)cbs * 4096);
          const int grid = prop.multiProcessorCount * occ;
          float best = 1e30f;
          for (int it = 0; it < 4; ++it) {
            cudaEventRecord(e0);
            copy_boxes<<<grid, 256>>>(src, dst, row_blocks, cbs, blocked, do_write);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
          }
          const double moved = (double)row_blocks * cbs * 4096 * (do_write ? 2 : 1);
          printf("%s  pitch %5d B  %-10s  %d blocks/SM: %7.1f GB/s\n", do_write ? "copy" : "read", cbs * 128,
                 blocked ? "K-blocked" : "row-major", occ, moved / best / 1e6);
        }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
