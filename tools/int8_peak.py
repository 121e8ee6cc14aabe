#!/usr/bin/env python
"""Measured INT8 tensor-core GEMM rate of this GPU (library GEMM: torch._int_mm -> cuBLASLt), the
denominator of the Ozaki split-integer feasibility estimate in DESIGN.md section 9.  Prints one JSON line.
Shapes: a large square GEMM (the achievable peak) and the shape the Gram would expose per trial
(M = 2(N+1) padded to 256, N = 16 weight columns x 2(N+1), K = T_d), batched as one tall GEMM."""
import json
import torch


def rate(m, n, k, reps=20):
    a = torch.randint(-127, 127, (m, k), dtype=torch.int8, device="cuda")
    b = torch.randint(-127, 127, (k, n), dtype=torch.int8, device="cuda")
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch._int_mm(a, b)
    e1.record()
    torch.cuda.synchronize()
    return 2.0 * m * n * k * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12


out = {"int8_tops_8192cube": rate(8192, 8192, 8192),
       "int8_tops_gram_shape_m256_n2080_k256_x64trials": rate(256 * 64, 2080, 256),
       "int8_tops_k256_square": rate(8192, 8192, 256)}
print(json.dumps(out))
