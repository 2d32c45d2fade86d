"""Byte-SIMD identities of the packed LDPC decoder (csrc/ldpc_packed_simd.cuh) checked on the CPU.

The header compiles for the host with NRB200_HOST_EMUL (plain C emulations of prmt / lop3 / vabsdiff4 / mad.lo); the checker
(tests/host/packed_simd_check.cc) compares cn_input / twomin / make_r, composed exactly as the kernel's cn_row composes them,
against the scalar definition of the reference's check-node update (nrLDPC_cnProc.h:388-877 on inputs formed as in
nrLDPC_bnProc.h:325): exhaustively per byte for the input stage, randomised whole rows for every row degree of BG1 / BG2."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_simd_identities(tmp_path):
    exe = str(tmp_path / "packed_simd_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "openairinterface5g_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "host", "packed_simd_check.cc")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "packed_simd_check OK" in out.stdout
