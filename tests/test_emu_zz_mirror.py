"""The GPU parity tests of the kernels that have not met a B200 yet (tests/test_gpu_zz_*.py) — run here, unchanged, on
CPU tensors over the emulated library (tests/emu_host.py) with their DEV constant switched to "cpu".  This checks the
*test code itself* (shapes, tolerances, oracle calls), so that the first hardware run fails only for hardware reasons."""
import pytest
import torch

from emu_host import emulated_host  # noqa: F401  (fixture)


def test_densify_gpu_tests_on_the_emulator(emulated_host, monkeypatch):
    import test_gpu_zz_densify as t
    monkeypatch.setattr(t, "DEV", "cpu")
    for path in t.GOLDEN:
        t.test_cuda_matches_reference_golden(path)
    for args in ((5_000, 1, 0, 2, None), (1, 2, 1, 2, 20), (255, 3, 2, 2, 20), (257, 3, 2, 2, None),
                 (3_001, 4, 1, 3, 20)):
        t.test_cuda_matches_oracle_on_random_models(*args)
    t.test_nothing_selected_and_everything_pruned()
    t.test_model_level_drop_in_matches_seeded_reference_flow()


def test_extract_gpu_tests_on_the_emulator(emulated_host, monkeypatch):
    import test_gpu_zz_extract as t
    monkeypatch.setattr(t, "DEV", "cpu")
    for name in ("a", "b", "c"):
        t.test_part_colours_match_reference_golden(name)
    t.test_epilogue_matches_oracle_on_a_large_image()


def test_regularizer_gpu_tests_on_the_emulator(emulated_host, monkeypatch):
    import test_gpu_zz_regularizers as t
    monkeypatch.setattr(t, "DEV", "cpu")
    t.test_fused_regularizers_match_the_reference_expressions(77, 123, False)
    t.test_fused_regularizers_match_the_reference_expressions(300, 400, True)
