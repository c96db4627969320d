"""The C++ host adapter (mcell_b200/host: reference-shaped BaseEvent / Molecule / Partition mirror above the
C ABI) compiled with g++ and run as a native program.  CPU tier: it builds, links against libmcx.so and fails
loudly without a device.  GPU tier: free-diffusion MSD/containment and A+B->C count identities through the
scheduler-style barrier loop."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "host", "test_host_adapter")


def _build():
    from mcell_b200 import build as b
    b.build()
    srcs = [os.path.join(ROOT, "tests", "host", "test_host_adapter.cpp"), os.path.join(ROOT, "mcell_b200", "host", "mcx_host.cpp")]
    deps = srcs + [os.path.join(ROOT, "mcell_b200", "host", "mcx_host.h"), os.path.join(ROOT, "include", "mcx.h")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return
    libdir = os.path.join(ROOT, "mcell_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", EXE] + srcs +
                   ["-L" + libdir, "-l:libmcx.so", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)


def test_count_buffer_text_formats(tmp_path):
    """CountBuffer .dat / .gdat writer of the host adapter: byte-identical to the reference's stream formatting."""
    exe = os.path.join(ROOT, "tests", "host", "test_count_buffer")
    srcs = [os.path.join(ROOT, "tests", "host", "test_count_buffer.cpp"), os.path.join(ROOT, "mcell_b200", "host", "mcx_host.cpp")]
    from mcell_b200 import build as b
    b.build()
    libdir = os.path.join(ROOT, "mcell_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe] + srcs +
                   ["-L" + libdir, "-l:libmcx.so", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "count buffer ok" in r.stdout


def test_viz_output_formats(tmp_path):
    """VizOutputWriter / GpuVizOutputEvent of the host adapter: ASCII (%.9g) and CellBlender binary v1/v2 molecule dumps,
    file naming and species grouping as in src4/viz_output_event.cpp:65-265."""
    exe = os.path.join(ROOT, "tests", "host", "test_viz_output")
    srcs = [os.path.join(ROOT, "tests", "host", "test_viz_output.cpp"), os.path.join(ROOT, "mcell_b200", "host", "mcx_host.cpp")]
    from mcell_b200 import build as b
    b.build()
    libdir = os.path.join(ROOT, "mcell_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe] + srcs +
                   ["-L" + libdir, "-l:libmcx.so", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "viz output ok" in r.stdout


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_adapter_builds_and_fails_loudly_without_device():
    _build()
    if _has_gpu():
        pytest.skip("device present: covered by the gpu test")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_adapter_runs_scheduler_loop_on_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "host adapter ok" in r.stdout
