// CTA-pair GEMM on f16mx operands (mx.cuh).  Filled in below.
#pragma once
#include "gemm.cuh"
#include "mx.cuh"
