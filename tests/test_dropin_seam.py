"""CPU tier for the drop-in surface: dropin/py_arkworks_bls12381 running on the host-emulated
kernels (tests/conftest.py::build_seam).  Includes full prove/verify against the reference's
golden proof bytes at N=8/16, and the UNMODIFIED reference on top of the drop-in when mounted."""
import importlib
import os
import sys

import pytest

import dropin_cases as dc
import shuffle_cases as sc


@pytest.fixture(scope="module")
def dropin(seam_lib):
    from curdleproofs_pie_b200 import runtime

    runtime._install_library_for_tests(seam_lib)
    mod = importlib.import_module("py_arkworks_bls12381")
    yield mod
    runtime._install_library_for_tests(None)


def test_surface_kats(dropin):
    dc.surface_kats(dropin)


def test_multiexp(dropin, cref):
    dc.multiexp_matches_oracle(dropin, cref, 24)


def test_deferred_decoding(dropin, cref):
    dc.deferred_decoding(dropin, cref)


def test_verify_verdicts_equal_reference_with_deferred_decoding(dropin):
    prev = dropin.defer_decoding(True)
    try:
        getattr(dropin, "_decoded", {}).clear()
        sc.check_verify_matches_golden(dropin, sc.load_case("shuffle_N8_seed1234.json"))
    finally:
        dropin.defer_decoding(prev)


@pytest.mark.parametrize("name", ["shuffle_N8_seed1234.json", "shuffle_N16_seed77.json"])
def test_prove_bytes_equal_reference(dropin, name):
    sc.check_prove_matches_golden(dropin, sc.load_case(name))


def test_verify_verdicts_equal_reference(dropin):
    sc.check_verify_matches_golden(dropin, sc.load_case("shuffle_N8_seed1234.json"))
