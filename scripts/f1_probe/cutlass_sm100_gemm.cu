// Compile probe only (see README.md in this directory): a CUTLASS 4.5 CollectiveBuilder GEMM for arch::Sm100,
// to check that the vendored headers produce tcgen05 / TMA / TMEM code offline.  Not product code, never run.
#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/gemm/dispatch_policy.hpp"
#include "cutlass/util/packed_stride.hpp"
using namespace cute;
using ElementA = cutlass::half_t; using LayoutA = cutlass::layout::RowMajor; constexpr int AlignmentA = 8;
using ElementB = cutlass::half_t; using LayoutB = cutlass::layout::ColumnMajor; constexpr int AlignmentB = 8;
using ElementC = cutlass::half_t; using LayoutC = cutlass::layout::RowMajor; constexpr int AlignmentC = 8;
using ElementAccumulator = float;
using ArchTag = cutlass::arch::Sm100;
using OperatorClass = cutlass::arch::OpClassTensorOp;
using MmaTileShape_MNK = Shape<_256,_128,_64>;
using ClusterShape_MNK = Shape<_2,_1,_1>;
using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    ArchTag, OperatorClass, MmaTileShape_MNK, ClusterShape_MNK,
    cutlass::epilogue::collective::EpilogueTileAuto,
    ElementAccumulator, ElementAccumulator,
    ElementC, LayoutC, AlignmentC, ElementC, LayoutC, AlignmentC,
    cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;
using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    ArchTag, OperatorClass, ElementA, LayoutA, AlignmentA, ElementB, LayoutB, AlignmentB, ElementAccumulator,
    MmaTileShape_MNK, ClusterShape_MNK,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
    cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;
using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int,int,int,int>, CollectiveMainloop, CollectiveEpilogue, void>;
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;
extern "C" int run_gemm(const void *A, const void *B, void *C, int M, int N, int K, void *ws, size_t ws_bytes, cudaStream_t s) {
  using StrideA = typename Gemm::GemmKernel::StrideA; using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC; using StrideD = typename Gemm::GemmKernel::StrideD;
  auto sa = cutlass::make_cute_packed_stride(StrideA{}, cute::make_shape(M, K, 1));
  auto sb = cutlass::make_cute_packed_stride(StrideB{}, cute::make_shape(N, K, 1));
  auto sc = cutlass::make_cute_packed_stride(StrideC{}, cute::make_shape(M, N, 1));
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, 1},
    {(const ElementA*)A, sa, (const ElementB*)B, sb}, {{1.f, 0.f}, (const ElementC*)C, sc, (ElementC*)C, sc}};
  Gemm gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) return -1;
  if (Gemm::get_workspace_size(args) > ws_bytes) return -2;
  if (gemm.initialize(args, ws, s) != cutlass::Status::kSuccess) return -3;
  return gemm.run(s) == cutlass::Status::kSuccess ? 0 : -4;
}
