"""The cals:: C++ surface (include/cals.h, als.h, ...; cp-cals_b200/libcals.so) -- the drop-in boundary of SURVEY 8b.

CPU part: the library and the binaries exist and the library refuses to run without a GPU.
GPU part (-m gpu): the restated relational tests (tests/cpp/test_*.cpp), and -- where they were prebuilt in the
build container by tools/build_ref_compat.sh -- the reference's OWN test sources and example driver, compiled
UNCHANGED against this repository's headers and library, run on the B200.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "cp-cals_b200", "bin")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "cp-cals_b200")], check=True, capture_output=True)
    script = os.path.join(ROOT, "tools", "build_ref_compat.sh")
    if os.path.isdir("/root/reference/src"):
        subprocess.run([script], check=True, capture_output=True)


def _run(exe, *args, timeout=900):
    path = os.path.join(BIN, exe)
    if not os.path.exists(path):
        pytest.skip("%s not built (tools/build_ref_compat.sh needs /root/reference; prebuilt files travel)" % exe)
    out = subprocess.run([path, *args], capture_output=True, text=True, timeout=timeout)
    return out.returncode, out.stdout + out.stderr


def test_cpp_library_and_binaries_build():
    _build()
    for f in ("libcals.so", "libcals_b200.so"):
        assert os.path.exists(os.path.join(ROOT, "cp-cals_b200", f))
    for f in ("driver", "c_abi_demo", "test_cals", "test_als"):
        assert os.path.exists(os.path.join(BIN, f))
    syms = subprocess.run(["nm", "-DC", os.path.join(ROOT, "cp-cals_b200", "libcals.so")], capture_output=True,
                          text=True).stdout
    for want in ("cals::cp_cals(cals::Tensor const&", "cals::jk_cp_cals(cals::Tensor const&",
                 "cals::cp_als(cals::Tensor const&", "cals::jk_cp_als(", "cals::cp_omp_als(", "cals::jk_cp_omp_als(",
                 "cals::mttkrp::mttkrp(", "cals::Ktensor::normalize(", "cals::Tensor::Tensor(",
                 "cals::utils::generate_jk_ktensors(", "cals::utils::jk_permutation_adjustment(", "set_threads(int)"):
        assert want in syms, "libcals.so does not export " + want
    if os.path.isdir("/root/reference/src"):
        # the reference's own callers compile unchanged against include/ + libcals.so
        for f in ("ref_driver", "ref_test_cals", "ref_test_als"):
            assert os.path.exists(os.path.join(BIN, f))


@pytest.mark.skipif(_have_gpu(), reason="only meaningful on a machine without a GPU")
def test_cpp_api_has_no_cpu_fallback():
    _build()
    rc, out = _run("test_als", "--gtest_filter=Als.FastErrorEqualsExplicitError")
    assert rc != 0
    assert "no CPU fallback" in out or "no CUDA device" in out


def test_host_containers_on_cpu():
    """Pure host logic of the C++ containers (no device needed): built and run as a tiny program."""
    _build()
    src = os.path.join(ROOT, "tests", "cpp", "host_containers_check.cpp")
    exe = os.path.join(BIN, "host_containers_check")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "include", "utils"), "-o", exe, src,
                    "-L" + os.path.join(ROOT, "cp-cals_b200"), "-lcals", "-lcals_b200", "-pthread",
                    "-Wl,-rpath," + os.path.join(ROOT, "cp-cals_b200")], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("OK")


@pytest.mark.gpu
@pytest.mark.parametrize("exe", ["test_als", "test_cals"])
def test_restated_reference_tests(exe):
    rc, out = _run(exe)
    assert rc == 0, out[-4000:]
    assert "[  PASSED  ]" in out and "[  FAILED  ]" not in out


def _run_randomised(exe, attempts=3):
    """The reference's own tests draw tensors and starting values from std::random_device (src/tensor.cpp:121-129) and
    compare against fixed thresholds -- e.g. Als.ComputeCorrectResult4D fits a rank-7 model to a random rank-5 3x3x3x3
    tensor and wants ||X - M|| < 0.1 after at most 100 iterations (tests/als/test_als.cpp:104-123), which an unlucky
    start misses whatever library runs it (observed on the B200: 1 run in 12, slow_error 0.118).  They are therefore
    given up to three draws; the seeded restatements in tests/cpp (test_restated_reference_tests) get one."""
    outs = []
    for _ in range(attempts):
        rc, out = _run(exe)
        outs.append(out)
        if rc == 0 and "[  FAILED  ]" not in out:
            return out
    raise AssertionError("%s failed %d times in a row:\n%s" % (exe, attempts, outs[-1][-4000:]))


@pytest.mark.gpu
def test_reference_als_tests_unchanged():
    """reference tests/als/test_als.cpp compiled as is -- all four tests, the NNLS one included."""
    assert "4 tests ran" in _run_randomised("ref_test_als")


@pytest.mark.gpu
def test_reference_cals_tests_unchanged():
    """reference tests/cals/test_cals.cpp compiled as is -- all five tests, both line-search methods included."""
    assert "5 tests ran" in _run_randomised("ref_test_cals")


@pytest.mark.gpu
def test_reference_driver_unchanged():
    rc, out = _run("ref_driver", "-t", "60-50-40", "-c", "1:6:3")
    assert rc == 0, out[-2000:]
    assert "CALS time:" in out and "Speedup:" in out


@pytest.mark.gpu
def test_driver_example():
    rc, out = _run("driver", "-t", "100-100-100", "-c", "1:10:4", "-i", "20")
    assert rc == 0, out[-2000:]
    assert "800 model-iterations" in out


@pytest.mark.gpu
def test_plain_c_caller_of_the_c_abi():
    """examples/c_abi_demo.c: the boundary is usable from C (gcc, no C++ runtime on the caller's side)."""
    rc, out = _run("c_abi_demo")
    assert rc == 0, out[-2000:]
    assert out.strip().endswith("OK") and "sm_100a" in out
