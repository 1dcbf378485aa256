// kungfu_b200 runtime — host reduction kernels of the CPU collective backend.
//
// Reference: `std_transform_2` (srcs/go/kungfu/base/op.cpp:57-93, 11 dtypes x {sum,min,max,prod}) and the AVX/F16C
// `float16_sum` (srcs/go/kungfu/base/f16.c:16-50). Here bf16 is first-class (the reference mis-maps it to fp16,
// srcs/cpp/include/kungfu/tensorflow/ops.h:21-22) and 16-bit floats accumulate through fp32.
#pragma once
#include <cstddef>
#include <cstdint>

namespace kfrt {

// codes shared with python (utils/dtypes.py) and the CUDA kernels (csrc/kernels/mesh_args.h)
enum DType : int { kF32 = 0, kBF16 = 1, kF16 = 2, kI32 = 3, kI64 = 4, kU8 = 5, kF64 = 6, kI8 = 7, kU32 = 8, kU64 = 9, kI16 = 10, kU16 = 11 };
enum RedOp : int { kSum = 0, kMin = 1, kMax = 2, kProd = 3 };

size_t dtype_size(int dt);
const char* dtype_name(int dt);
const char* op_name(int op);

// dst[i] = x[i] (op) y[i]; dst may alias x or y.
void transform2(void* dst, const void* x, const void* y, size_t count, int dtype, int op);

float half_to_float(uint16_t h);
uint16_t float_to_half(float f);
float bf16_to_float(uint16_t h);
uint16_t float_to_bf16(float f);

} // namespace kfrt
