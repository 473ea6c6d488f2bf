// Stub used ONLY to compile the reference's *_gpu.cu files (oracle/_ref) without torch
// headers: their *_gpu.h headers declare pybind wrappers taking at::Tensor by value; a
// forward declaration is enough for a declaration that is never defined or called here.
#pragma once
namespace at { class Tensor; }
