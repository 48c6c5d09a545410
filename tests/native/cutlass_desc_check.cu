// TEST INFRASTRUCTURE (tests/test_cutlass_descriptors.py): prints the tcgen05 descriptors CUTLASS/CuTe itself builds for the
// shared-memory layout and instruction shape of candmc_b200/csrc/gemm_f32.cu next to the ones the product's encoders
// (common.cuh) produce.  Host-only program: CuTe cannot take a shared-memory address on the host, so the start address field
// is 0 on both sides; everything else (stride / leading byte offsets, version, layout type, the instruction descriptor, and
// the element offsets that the kernel's address arithmetic assumes) is compared.
#include <cstdint>
#include <cstdio>

#include <cute/tensor.hpp>
#include <cute/arch/mma_sm100_desc.hpp>
#include <cute/atom/mma_traits_sm100.hpp>

#include "common.cuh"

using namespace cute;

int main() {
  using T = cutlass::tfloat32_t;
  // K-major 128 x 32 tile of TF32 under the 128-byte swizzle: what a TMA box {32 floats, 128 rows} with SWIZZLE_128B writes
  auto layout = tile_to_shape(UMMA::Layout_K_SW128_Atom<T>{}, Shape<_128, _32>{});
  alignas(1024) static T fake[128 * 32];
  Tensor s = make_tensor(make_smem_ptr(fake), layout);
  Tensor s0 = local_tile(s, Shape<_128, _8>{}, make_coord(0, 0));   // one UMMA instruction: M = 128, K = 8
  UMMA::SmemDescriptor d = UMMA::make_umma_desc<UMMA::Major::K>(s0);
  auto id = UMMA::make_instr_desc<T, T, float, 128, 128, UMMA::Major::K, UMMA::Major::K>();
  auto plain = layout.layout_b();   // the layout without the swizzle functor: where (row, k) sits before the XOR
  printf("{\"cutlass_smem_desc\": \"%016llx\", \"candmc_smem_desc\": \"%016llx\", \"cutlass_idesc\": \"%08x\", \"candmc_idesc\": \"%08x\", "
         "\"elem_offset_row1\": %d, \"elem_offset_row8\": %d, \"elem_offset_k8\": %d, \"elem_offset_k24\": %d}\n",
         (unsigned long long)d.desc_, (unsigned long long)candmc::umma_desc_kmajor_sw128(0), (unsigned)id.desc_,
         (unsigned)candmc::umma_idesc_tf32(128, 128), (int)plain(1, 0), (int)plain(8, 0), (int)plain(0, 8), (int)plain(0, 24));
  // the swizzle itself: CuTe's composed layout against the XOR the simulator's TMA and UMMA emulation apply to byte addresses
  // of a 1024-byte-aligned tile (16-byte chunk index ^= 128-byte line index mod 8)
  int mismatches = 0;
  for (int row = 0; row < 128; ++row)
    for (int k = 0; k < 32; ++k) {
      const uint32_t cute_bytes = (uint32_t)((const char*)&s(row, k) - (const char*)&s(0, 0));   // `fake` is 1024-byte aligned
      uint32_t addr = (uint32_t)row * 128u + (uint32_t)k * 4u;
      addr ^= ((addr >> 7) & 7u) << 4;
      if (addr != cute_bytes) ++mismatches;
    }
  printf("{\"swizzle_mismatches\": %d}\n", mismatches);
  return 0;
}
