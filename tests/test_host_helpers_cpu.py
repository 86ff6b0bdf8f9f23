"""CPU: host-side sizing helpers of the round-2 kernels (no GPU): persistent-grid balancing and k-split selection."""
import ctypes as C

import numpy as np


def test_balanced_grid_never_adds_a_wave_and_frees_sms():
    from serl_b200 import _lib as L
    lib = L.load()
    lib.serl_balanced_grid.argtypes = [C.c_int, C.c_int]
    lib.serl_balanced_grid.restype = C.c_int
    for sms in (148, 132, 64):
        for items in list(range(1, 700)) + [1024, 2048, 4096]:
            g = lib.serl_balanced_grid(items, sms)
            full = min(items, sms)
            assert 1 <= g <= full
            assert -(-items // g) == -(-items // full)                   # same makespan in items per CTA
            if items > sms:
                assert g == -(-items // (-(-items // sms)))              # the smallest such grid
    assert lib.serl_balanced_grid(512, 148) == 128 and lib.serl_balanced_grid(256, 148) == 128 and lib.serl_balanced_grid(2048, 148) == 147
    assert lib.serl_balanced_grid(64, 148) == 64


def test_tgemm_splits_cover_k_with_whole_blocks_and_no_empty_split():
    from serl_b200.ops import tgemm_splits
    for K in (256, 580, 4096, 1000, 33):
        for want in (1, 2, 6, 12, 22, 24, 32):
            S = tgemm_splits(K, want)
            assert 1 <= S <= want
            kc = -(-(-(-K // S)) // 32) * 32
            assert kc % 32 == 0 and -(-K // kc) == S and (S - 1) * kc < K
    assert tgemm_splits(4096, 12) == 12 and tgemm_splits(4096, 24) == 22
