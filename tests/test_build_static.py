"""Static checks of the sm_100a build that need no GPU: the hot kernels must not spill to local memory
and must keep the occupancy their launch bounds promise (ptxas -v logs written by the Makefile), and the
FMA-only constant division the kernels use must equal IEEE division (exhaustive over every float
mantissa for a sample of divisors; scripts/check_const_div.c)."""
import glob
import os
import re
import subprocess

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "cuda_flow3d_b200", "csrc")


def _ptxas_entries():
    out = {}
    for path in glob.glob(os.path.join(CSRC, "*.ptxas.log")):
        name = None
        for line in open(path):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                name = m.group(1)
                out[name] = {"spill": 0, "regs": None}
                continue
            if name is None:
                continue
            m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                out[name]["spill"] = int(m.group(1)) + int(m.group(2))
            m = re.search(r"Used (\d+) registers", line)
            if m:
                out[name]["regs"] = int(m.group(1))
    return out


def test_built_for_sm100a_without_spills():
    e = _ptxas_entries()
    if not e:
        pytest.skip("no ptxas logs (library built elsewhere)")
    hot = [k for k in e if re.search(r"sweep_kernel|phi_ksi_kernel|median5_march|resample|warp_derivatives|update_norm", k)]
    assert hot, "hot kernels missing from the ptxas logs"
    for k in hot:
        assert e[k]["spill"] == 0, "%s spills %d bytes" % (k, e[k]["spill"])
    # launch bounds: VEC=4 sweeps run 2 CTAs of 128 threads per SM (<= 255 regs), VEC=2 sweeps 4 CTAs (<= 128),
    # the phi kernel at VEC=2 five CTAs (<= 102)
    for k, v in e.items():
        if "sweep_kernelILi4" in k:
            assert v["regs"] <= 255
        if "sweep_kernelILi2" in k:
            assert v["regs"] <= 128
        if "phi_ksi_kernelILi2ELb0" in k:
            assert v["regs"] <= 102


def test_constant_division_sequence_is_exact(tmp_path):
    exe = str(tmp_path / "check_const_div")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe,
                           os.path.join(ROOT, "scripts", "check_const_div.c"), "-lm"])
    r = subprocess.run([exe, "12", "7"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches=0" in r.stdout


def test_pdl_chain_kernels_wait_before_touching_global_memory():
    """Kernels that solve_level launches as programmatic dependents (common.cuh: launch_chain_kernel) may become
    resident while their predecessor still runs: each must execute griddepcontrol.wait (SASS: ACQBULK) before
    its first global load, store or prefetch.  Checked on the SASS of every instance of the two chain kernels."""
    obj = os.path.join(CSRC, "kernels_solve.o")
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not (os.path.exists(obj) and os.path.exists(cuobjdump)):
        pytest.skip("no object file / cuobjdump")
    sass = subprocess.run([cuobjdump, "-sass", obj], capture_output=True, text=True, timeout=600).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    checked = 0
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if not re.search(r"sweep_kernel|phi_ksi_kernel", name):
            continue
        ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)
        assert "ACQBULK" in ops and "PREEXIT" in ops, name
        first_wait = ops.index("ACQBULK")
        mem = [i for i, o in enumerate(ops) if re.match(r"(LDG|STG|LD\b|ST\b|CCTL|ATOM|RED|LDGSTS|UTMA)", o)]
        assert mem and min(mem) > first_wait, "%s touches global memory before ACQBULK" % name
        checked += 1
    assert checked >= 12
