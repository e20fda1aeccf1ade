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
