#!/usr/bin/env python3
"""Exhaustive search for exact single-multiply forms of Compress_d on canonical inputs (x < q):
floor(((x << d) + c) * M / 2^32) mod 2^d == floor((2^d x + 1664) / q) mod 2^d for ALL x in [0, q).
Used by compress_canon<D> in csrc/mlkem_device.cuh."""
q = 3329
for d in (1, 4, 5, 10, 11):
    ok = [(c, M) for M in range(1290160, 1290176) for c in range(1655, 1672)
          if all(((((x << d) + c) * M) >> 32) & ((1 << d) - 1) == ((x << d) + 1664) // q % (1 << d) for x in range(q))]
    print(d, ok)
