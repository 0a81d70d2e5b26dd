"""The C++ adapter (reindexer_b200/host/gpu_bruteforce.h) compiled against the reference's own headers: one template drives it and
hnswlib::BruteforceSearch through the calls HnswIndexBase<Map> makes and diffs the results (tests/cpp/dropin_check.cc)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "dropin_check")
BIN_HNSW = os.path.join(ROOT, "tests", "cpp", "_build", "dropin_hnsw_check")
BIN_IVF = os.path.join(ROOT, "tests", "cpp", "_build", "dropin_ivf_check")
BIN_FT = os.path.join(ROOT, "tests", "cpp", "_build", "dropin_ft_check")


def test_adapter_compiles_against_reference_headers():
    if not os.path.isdir("/root/reference/cpp_src"):
        pytest.skip("reference tree not present on this box (the prebuilt binary is used by the gpu test)")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "port"])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    assert os.path.exists(BIN) and os.path.exists(BIN_HNSW) and os.path.exists(BIN_FT)


@pytest.mark.gpu
def test_adapter_matches_reference_map_on_gpu():
    if not os.path.exists(BIN):
        pytest.skip("tests/cpp/_build/dropin_check was not built (needs /root/reference at build time)")
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("MATCH") == 4 and "MISMATCH" not in out.stdout


@pytest.mark.gpu
def test_hnsw_adapter_matches_reference_map_on_gpu():
    """reindexer_b200/host/gpu_hnsw.h vs hnswlib::HierarchicalNSW<None>: same inserter, search on the device (tests/cpp/dropin_hnsw_check.cc)"""
    if not os.path.exists(BIN_HNSW):
        pytest.skip("tests/cpp/_build/dropin_hnsw_check was not built (needs /root/reference at build time)")
    out = subprocess.run([BIN_HNSW], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("MATCH") == 3 and "MISMATCH" not in out.stdout


@pytest.mark.gpu
def test_ft_merge_adapter_matches_reference_merger_on_gpu():
    """reindexer_b200/host/gpu_ft_merge.h vs ft::Merger::Merge on QueryMergeData built from the reference's own containers
    (tests/cpp/dropin_ft_check.cc)"""
    if not os.path.exists(BIN_FT):
        pytest.skip("tests/cpp/_build/dropin_ft_check was not built (needs /root/reference at build time)")
    out = subprocess.run([BIN_FT], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "MISMATCH" not in out.stdout and out.stdout.count("MATCH") >= 4


@pytest.mark.gpu
def test_ivf_adapter_matches_reference_faiss_on_gpu():
    """reindexer_b200/host/gpu_ivf.h vs the reference's vendored faiss::IndexIVFFlat driven like IvfIndex (tests/cpp/dropin_ivf_check.cc):
    knn + range searches between bursts of upserts and deletes, the device lists patched in place (one import)"""
    if not os.path.exists(BIN_IVF):
        pytest.skip("tests/cpp/_build/dropin_ivf_check was not built (needs /root/reference at build time)")
    out = subprocess.run([BIN_IVF], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "MISMATCH" not in out.stdout and out.stdout.count("MATCH") == 3
