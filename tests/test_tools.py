"""CPU: the profile tooling reads the committed ncu launch list (profiles/) and reproduces the committed per-class JSON."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ncu_class_summary_reproduces_the_committed_numbers(tmp_path):
    tool = os.path.join(ROOT, "multi-feature-vit_b200", "tools", "ncu_classes.py")
    src = os.path.join(ROOT, "profiles", "r01_ncu_launches_b32_v3.csv")
    out = os.path.join(str(tmp_path), "classes.json")
    r = subprocess.run([sys.executable, tool, src, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = json.load(open(out))
    want = json.load(open(os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")))
    assert got["classes"] == want["classes"] and got["total_us"] == want["total_us"]
    cls = got["classes"]
    # one step: 12 blocks x 4 forward GEMMs + patch embedding, 4 dgrads and 4 weight gradients per block (+ patch embedding)
    assert cls["gemm_fwd"]["launches_per_step"] == 49 and cls["gemm_dgrad"]["launches_per_step"] == 48
    assert cls["gemm_wgrad"]["launches_per_step"] == 49 and cls["attn_fwd"]["launches_per_step"] == 12
    assert cls["ln_fwd"]["launches_per_step"] == cls["ln_bwd"]["launches_per_step"] == 25
    assert abs(sum(c["share"] for c in cls.values()) - 1.0) < 1e-3


def test_ncu_class_summary_of_the_round_2_step(tmp_path):
    """Round 2: LayerNorm forward fused into the proj / fc2 GEMMs (1 standalone launch left), the patch embedding is its
    own tcgen05 kernel (48 forward GEMM launches), the fusion runs as batched stage kernels."""
    tool = os.path.join(ROOT, "multi-feature-vit_b200", "tools", "ncu_classes.py")
    src = os.path.join(ROOT, "profiles", "r02_ncu_launches_b32.csv")
    out = os.path.join(str(tmp_path), "classes.json")
    r = subprocess.run([sys.executable, tool, src, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = json.load(open(out))
    want = json.load(open(os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")))
    assert got["classes"] == want["classes"] and got["total_us"] == want["total_us"]
    cls = got["classes"]
    assert cls["gemm_fwd"]["launches_per_step"] == 48 and cls["gemm_dgrad"]["launches_per_step"] == 48
    assert cls["gemm_wgrad"]["launches_per_step"] == 49
    assert cls["ln_fwd"]["launches_per_step"] == 1 and cls["ln_bwd"]["launches_per_step"] == 25
    assert "patch_embed_kernel" in r.stdout and "fus2_stream_fwd_kernel" in r.stdout


def test_clock_sampler_counts_only_rows_inside_the_timed_window():
    """bench.py's nvidia-smi sampler is started before the warm-up and reports the rows that arrived between begin() and
    end() (the timed regions); with none inside it says so and falls back to the whole run."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")

    class _Proc:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    row = lambda mhz, cap: "%d, 1965, 700.0, Not Active, Not Active, Not Active, %s" % (mhz, cap)
    s = bench.ClockSampler(0)
    s.proc = _Proc()
    s.rows = [(10.0, row(1200, "Not Active")), (20.0, row(1965, "Not Active")), (21.0, row(1950, "Active")),
              (22.0, row(1965, "Not Active")), (30.0, row(900, "Not Active"))]
    s.t_begin, s.t_end = 19.5, 22.5
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
    s2 = bench.ClockSampler(0)
    s2.proc = _Proc()
    s2.rows = [(10.0, row(1200, "Not Active"))]
    s2.t_begin, s2.t_end = 19.5, 22.5
    out2 = s2.stop()
    assert out2["samples"] == 0 and out2["samples_whole_run"] == 1 and out2["sm_mhz"] == 1200.0
