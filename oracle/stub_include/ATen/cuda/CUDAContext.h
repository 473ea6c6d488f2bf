// Stub, see ../../torch/serialize/tensor.h
#pragma once
#include <cuda_runtime_api.h>
namespace at { class Tensor; }
