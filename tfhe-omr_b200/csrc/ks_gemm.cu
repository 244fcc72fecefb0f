// ks_gemm.cu — the LWE key switch (detector.rs:560-563) as an exact int8 tensor-core GEMM on sm_100a.
//   out[m][col] = SUM_{i<1024, j<27} d_ij(m) * KSK[i][j][col],  d in {-1,0,1}
// is C[M x N] = A[M x K] * B[K x N] with A = balanced base-2 digits of the extracted LWE mask (int8, K = 27 648), B = the key
// split into four balanced base-256 limbs (int8 in [-128,127], N = 671 * 4 padded to 2 688), int32 accumulation: every
// partial sum is bounded by 27 648 * 128 < 2^31, so the tensor-core result is exact and the limbs recombine to the same
// integer the CUDA-core kernel (keyswitch_kernel) accumulates.  The GEMM itself is a CUTLASS 4 / CuTe collective for
// Sm100 (tcgen05.mma kind::i8, accumulators in TMEM, operands staged by TMA) instantiated here; the digit expansion and the
// limb recombination + modulus switch are the kernels in kernels.cuh (ks_digits_kernel, ks_combine_kernel).
// Compiled only when the CUTLASS headers are found at build time (build.py); otherwise omr::ks_gemm_i8 reports "unavailable"
// and the library keeps using keyswitch_kernel.
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace omr {
int ks_gemm_i8(const int8_t* A, const int8_t* B, int32_t* C, int M, int N, int K, void* workspace, size_t workspace_bytes, cudaStream_t s);
size_t ks_gemm_workspace(int M, int N, int K);
bool ks_gemm_available();
}  // namespace omr

#ifdef OMR_HAVE_CUTLASS
#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/numeric_types.h"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/util/packed_stride.hpp"

namespace {
using namespace cute;
using ElementA = int8_t;  using LayoutA = cutlass::layout::RowMajor;    constexpr int AlignA = 16;   // digits [M][K]
using ElementB = int8_t;  using LayoutB = cutlass::layout::ColumnMajor; constexpr int AlignB = 16;   // key limbs, K contiguous per column
using ElementC = int32_t; using LayoutC = cutlass::layout::RowMajor;    constexpr int AlignC = 4;
using ElementAcc = int32_t;
using ArchTag = cutlass::arch::Sm100;
using OpClass = cutlass::arch::OpClassTensorOp;
using MmaTile = Shape<_128, _128, _128>;
using ClusterShape = Shape<_1, _1, _1>;
using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    ArchTag, OpClass, MmaTile, ClusterShape, cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementAcc,
    ElementC, LayoutC, AlignC, ElementC, LayoutC, AlignC, cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;
using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    ArchTag, OpClass, ElementA, LayoutA, AlignA, ElementB, LayoutB, AlignB, ElementAcc, MmaTile, ClusterShape,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
    cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;
using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>;
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;

typename Gemm::Arguments make_args(const int8_t* A, const int8_t* B, int32_t* C, int M, int N, int K) {
    using StrideA = typename Gemm::GemmKernel::StrideA; using StrideB = typename Gemm::GemmKernel::StrideB;
    using StrideC = typename Gemm::GemmKernel::StrideC; using StrideD = typename Gemm::GemmKernel::StrideD;
    const StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, {M, K, 1});
    const StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, {N, K, 1});
    const StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, {M, N, 1});
    const StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, {M, N, 1});
    return typename Gemm::Arguments{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, 1}, {A, sa, B, sb}, {{1, 0}, C, sc, C, sd}};
}
}  // namespace

namespace omr {
bool ks_gemm_available() { return true; }
size_t ks_gemm_workspace(int M, int N, int K) { return Gemm::get_workspace_size(make_args(nullptr, nullptr, nullptr, M, N, K)); }
int ks_gemm_i8(const int8_t* A, const int8_t* B, int32_t* C, int M, int N, int K, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    auto args = make_args(A, B, C, M, N, K);
    Gemm gemm;
    if (gemm.can_implement(args) != cutlass::Status::kSuccess) return 1;
    if (Gemm::get_workspace_size(args) > workspace_bytes) return 2;
    if (gemm.initialize(args, workspace, s) != cutlass::Status::kSuccess) return 3;
    return gemm.run(s) == cutlass::Status::kSuccess ? 0 : 4;
}
}  // namespace omr
#else
namespace omr {
bool ks_gemm_available() { return false; }
size_t ks_gemm_workspace(int, int, int) { return 0; }
int ks_gemm_i8(const int8_t*, const int8_t*, int32_t*, int, int, int, void*, size_t, cudaStream_t) { return -1; }
}  // namespace omr
#endif
