// umma_unaligned_probe.cu -- hardware probe (sm_100a): can a K-major SWIZZLE_128B tcgen05.mma operand START at a row that
// is not a multiple of 8 (start address not 1024-byte aligned), and with a stride between 8-row groups that is not a
// multiple of 1024 B?  A conv tile whose three horizontal taps read ONE halo box (start + s * 128 B) needs exactly that.
//
// A: 512 rows x 64 bf16 in shared memory, written the way TMA writes a 128B-swizzled box anchored at a 1024-aligned
// address: row p at p * 128, 16-byte chunk c at chunk (c ^ (p & 7)).  A[p][k] = value(p, k).  B = 64 x 64 identity, so
// D[m][n] = A[row(m)][n] names the row the tensor core actually read for accumulator row m.
//   expected: row(m) = off + (m / 8) * (SBO / 128) + (m % 8)
// Variants: off in {0,1,2,3,5,8,9,10,11}, SBO in {1024, 1280}, descriptor base-offset field 0 or (start >> 7) & 7.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -I unet-pytorch_b200/csrc -o /tmp/probe scripts/umma_unaligned_probe.cu && /tmp/probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "b2u_ptx.cuh"

using namespace b2u;

__device__ __host__ inline float value(int p, int k) { return static_cast<float>((p * 5 + k * 3) % 61) - 30.f; }

struct Variant { int off, sbo, use_base; };

__global__ void __launch_bounds__(128, 1) probe(const Variant* vars, int nvar, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  constexpr int kRows = 512;
  const uint32_t sA = base, sB = base + kRows * 128, bar = sB + 64 * 128, slot = bar + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A rows, swizzled
  for (int i = tid; i < kRows * 8; i += 128) {
    const int p = i >> 3, c = i & 7;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16(value(p, c * 8 + e));
    *reinterpret_cast<uint4*>(gen + p * 128 + ((c ^ (p & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
  }
  // B = identity [n][k], swizzled the same way
  for (int i = tid; i < 64 * 8; i += 128) {
    const int n = i >> 3, c = i & 7;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16(c * 8 + e == n ? 1.f : 0.f);
    *reinterpret_cast<uint4*>(gen + kRows * 128 + n * 128 + ((c ^ (n & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
  uint32_t phase = 0;
  for (int v = 0; v < nvar; ++v) {
    if (tid == 0) {
      const uint32_t start = sA + vars[v].off * 128;
      const uint64_t boff = vars[v].use_base ? static_cast<uint64_t>((start >> 7) & 7u) << 49 : 0ull;
      const uint64_t ad = umma_smem_desc(start, 16, vars[v].sbo, 2u) | boff;
      const uint64_t bd = umma_smem_desc(sB, 16, 1024, 2u);
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem, ad + ((k * 32) >> 4), bd + ((k * 32) >> 4), idesc, k ? 1u : 0u);
      tc_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
    uint32_t r[64];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), r);
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 32, r + 32);
    tmem_ld_wait();
    float* o = out + (static_cast<size_t>(v) * 128 + warp * 32 + lane) * 64;
    for (int n = 0; n < 64; ++n) o[n] = __uint_as_float(r[n]);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<Variant> vars;
  const int offs[] = {0, 1, 2, 3, 5, 8, 9, 10, 11, 21};
  for (int sbo : {1024, 1280, 2304})
    for (int off : offs)
      for (int ub : {0, 1}) vars.push_back({off, sbo, ub});
  Variant* dv; float* dout;
  cudaMalloc(&dv, vars.size() * sizeof(Variant));
  cudaMalloc(&dout, vars.size() * 128 * 64 * sizeof(float));
  cudaMemcpy(dv, vars.data(), vars.size() * sizeof(Variant), cudaMemcpyHostToDevice);
  const int smem = 512 * 128 + 64 * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dv, static_cast<int>(vars.size()), dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> out(vars.size() * 128 * 64);
  cudaMemcpy(out.data(), dout, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
  for (size_t v = 0; v < vars.size(); ++v) {
    int bad_rows = 0, first_bad = -1;
    for (int m = 0; m < 128; ++m) {
      const int row = vars[v].off + (m / 8) * (vars[v].sbo / 128) + (m % 8);
      bool ok = true;
      for (int n = 0; n < 64; ++n) ok = ok && out[(v * 128 + m) * 64 + n] == value(row, n);
      if (!ok) { ++bad_rows; if (first_bad < 0) first_bad = m; }
    }
    printf("off %2d  SBO %4d  base_offset field %s : %s", vars[v].off, vars[v].sbo, vars[v].use_base ? "(start>>7)&7" : "0           ",
           bad_rows == 0 ? "rows as expected" : "MISMATCH");
    if (bad_rows) printf(" (%d of 128 rows wrong, first m = %d)", bad_rows, first_bad);
    printf("\n");
  }
  return 0;
}
