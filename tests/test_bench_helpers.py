"""bench.py plumbing that has to hold on a GPU box we cannot rehearse on: byte model, counter differencing, the clock sampler (with a
fake nvidia-smi), argument handling of the reference arm under torchrun.  No GPU needed."""
import importlib.util
import json
import os
import stat
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_follow_the_survey_model():
    b = _bench()
    c = dict(paths=1000.0, extend_rays=1600.0, shade_invocations=1600.0, shadow_rays=1300.0)
    ab = b.algorithmic_bytes(c)
    assert ab["raygen"] == 68 * 1000 and ab["extend"] == 44 * 1600 and ab["resolve"] == 48 * 1000
    assert ab["shade"] == 156 * 1600 + 52 * 1300 and ab["connect"] == 76 * 1300 + 8 * 1600
    assert ab["total"] == 116 * 1000 + 224 * 1600 + 128 * 1300
    # the per-kernel figures add up to the total of the streaming model minus its 16 B/segment "classify + sort" row (queue indices: counted in the total only)
    assert abs(sum(ab[k] for k in ("raygen", "extend", "shade", "connect", "resolve")) - (ab["total"] - 16 * 1600)) < 1e-6


def test_counter_differencing():
    b = _bench()
    keys = ("paths", "extend_rays", "shade_invocations", "surface_hits", "misses", "shadow_rays", "medium_events", "kernel_launches")
    a = {k: i for i, k in enumerate(keys)}; z = {k: 3 * i + 5 for i, k in enumerate(keys)}
    d = b.diff_counters(a, z)
    assert d == {k: 2 * i + 5 for i, k in enumerate(keys)}


def test_clock_sampler_with_a_fake_nvidia_smi(tmp_path, monkeypatch):
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nwhile true; do echo '0, 1965, 1965, 703.2, 0x0000000000000004, Not Active, Not Active, Not Active, Active'; sleep 0.05; done\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    b = _bench()
    s = b.ClockSampler(0); s.start(); time.sleep(0.3); s.mark_begin(); time.sleep(0.3); s.mark_end(); r = s.stop()
    assert r["sm_mhz"] == 1965.0 and r["sm_max_mhz"] == 1965.0 and r["samples"] >= 2 and r["window"] == "timed region"
    assert r["reasons"] == ["sw_power_cap"]                                          # kept and noted, not a rejection reason
    s = b.ClockSampler(0); s.start(); time.sleep(0.3); s.mark_begin(); s.mark_end(); r = s.stop()   # a timed region shorter than the sampling period
    assert r["samples"] >= 2 and r["window"].startswith("warm-up + timed steps")
    monkeypatch.setenv("PATH", "/nonexistent")
    s = b.ClockSampler(0); s.start(); s.mark_begin(); s.mark_end(); r = s.stop()
    assert r["sm_mhz"] is None and r["reasons"] == ["nvidia-smi unavailable"]


def test_reference_arm_prints_the_contract_line_and_idle_ranks_exit_quietly():
    """`bench.py --impl reference` = the CPU oracle arm: rank 0 prints ONE JSON line with impl/cpu_baseline/e2e, other ranks print nothing."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_post_row_blocks_partition_the_image():
    import bench
    for H in (2160, 1080, 389):
        for world in (1, 2, 3, 4, 8):
            blocks = [bench.post_row_block(H, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:])) and all(b[1] > b[0] for b in blocks)
            assert all(b[0] % 8 == 0 for b in blocks)
