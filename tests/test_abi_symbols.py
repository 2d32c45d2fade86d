"""The C-ABI library loads and exports every symbol include/*.h declares (no compute call: works without a GPU)."""
import numpy as np
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_\s\*]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", txt, flags=re.M)
    return sorted(set(n for n in names if not n.startswith("check_crc")))


def test_ldpc_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "openairinterface5g_b200", "libldpc_b200.so"))
    names = _declared("nrb200_ldpc.h") + _declared("nrb200_slot.h") + _declared("nrb200_rfsim.h") + _declared("nrb200_prach.h")
    assert {"LDPCinit", "LDPCshutdown", "LDPCdecoder", "LDPCencoder", "nrb200_sch_slot_rx_dev", "nrb200_pdsch_slot_tx_dev", "nrb200_set_device",
            "nrb200_rfsim_rx_add_input_dev", "nrb200_rfsim_rx_add_input_host", "nrb200_rx_nr_prach_dev", "nrb200_rx_nr_prach_host"} <= set(names)
    for n in names:
        assert hasattr(lib, n), n


def test_dfts_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "openairinterface5g_b200", "libdfts_b200.so"))
    names = _declared("nrb200_dfts.h")
    assert {"dft", "idft", "dfts_autoinit"} <= set(names)
    for n in names:
        assert hasattr(lib, n), n


def test_only_abi_symbols_are_exported():
    """-fvisibility=hidden + -Bsymbolic: ldpctest loads two LDPC libraries RTLD_GLOBAL in one process."""
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "openairinterface5g_b200", "libldpc_b200.so")], text=True)
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(s.startswith(("LDPC", "nrb200_")) for s in syms), syms
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "openairinterface5g_b200", "libdfts_b200.so")], text=True)
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(s in ("dft", "idft", "dfts_autoinit") or s.startswith("nrb200_") for s in syms), syms


def test_struct_layouts_match_reference_abi():
    from openairinterface5g_b200 import ldpc
    assert ctypes.sizeof(ldpc.DecParams) == 40          # t_nrLDPC_dec_params (nrLDPC_types.h:84-97) on x86-64
    assert ldpc.DecParams.E.offset == 12 and ldpc.DecParams.check_crc.offset == 24
    assert ctypes.sizeof(ldpc.TimeStats) == 80 and ctypes.sizeof(ldpc.LdpcTimeStats) == 880
    assert ldpc.EncParams.Zc.offset == 56 and ldpc.EncParams.K.offset == 88


def test_pusch_num_llr_matches_oracle_bookkeeping(oracle):
    """nrb200_pusch_num_llr is host arithmetic (no GPU): G = sum over symbols of get_nb_re_pusch * Qm (nr_ulsch_demodulation.c:416-432, :1659-1664)."""
    from oracle.bindings import PuschParms
    from openairinterface5g_b200.ldpc import LdpcLib, PuschRxDesc
    lib = LdpcLib()
    for N, nb_rx, rb_size, Qm, dpos, dtype_, cdm, start, nsym in ((4096, 4, 273, 6, 1 << 2, 0, 2, 0, 14), (2048, 2, 50, 4, (1 << 2) | (1 << 11), 0, 1, 0, 14),
                                                                  (1024, 2, 32, 8, 1 << 2, 1, 2, 1, 12), (1024, 1, 52, 2, 1 << 3, 1, 1, 2, 10)):
        P = PuschParms(N, nb_rx, 0, 0, rb_size, N - 6 * rb_size, Qm, dpos, dtype_, cdm)
        want = sum(oracle.pusch_nb_re(P, s) for s in range(start, start + nsym)) * Qm
        for shift in (5, 0xFFFFFFFF):
            d = PuschRxDesc(N, nb_rx, 0, 0, rb_size, N - 6 * rb_size, Qm, start, nsym, dpos, dtype_, cdm, shift, 0, 0, 0, 0, 0)
            assert lib.pusch_num_llr(d) == want
    bad = PuschRxDesc(4096, 4, 0, 0, 273, 2458, 5, 0, 14, 4, 0, 2, 0, 0, 0, 0, 0, 0)       # Qm = 5
    assert lib.pusch_num_llr(bad) == 0


def test_pusch_dmrs_pilots_match_oracle(oracle):
    """nrb200_pusch_dmrs_pilots_host is host arithmetic (nr_gold_pusch + nr_pusch_dmrs_rx): no GPU needed."""
    from oracle.bindings import ChestParms
    from openairinterface5g_b200.ldpc import LdpcLib, PuschChestDesc
    lib = LdpcLib()
    for N, slot, symbol, port, rb_start, rb_size, scid, nid in ((4096, 1, 2, 0, 0, 273, 0, 77), (2048, 19, 11, 1, 30, 76, 1, 65535), (1024, 0, 2, 2, 0, 52, 1, 0),
                                                                (1024, 5, 0, 3, 20, 32, 0, 300)):
        P = ChestParms(N, 2, slot, symbol, port, rb_start, 0, rb_size, N - 6 * 52, scid, nid)
        d = PuschChestDesc(N, 2, slot, symbol, port, rb_start, 0, rb_size, N - 6 * 52, scid, nid, 14 * N, 14 * N, 1)
        assert np.array_equal(lib.pusch_dmrs_pilots(d), oracle.pusch_dmrs_pilots(P))


def test_offload_flavour_exports_the_loader_symbols():
    """libldpc_b200_t2.so (offload calling convention) loads next to libldpc_b200.so and exports exactly the four names the loader dlsym()s."""
    from openairinterface5g_b200.ldpc import OffloadLdpcLib
    lib = OffloadLdpcLib(init=False).lib
    for name in ("LDPCinit", "LDPCshutdown", "LDPCdecoder", "LDPCencoder"):
        assert getattr(lib, name) is not None


def test_packed_decoder_schedule_covers_every_item_and_balances():
    """Host arithmetic of the packed decoder's work lists (nrb200_ldpc_packed_schedule_info, no GPU): every (row, chunk) and (column, chunk) is
    owned by exactly one list for every lifting size the packed kernel serves, and the headline configurations balance to a few per cent."""
    import numpy as np
    lib = ctypes.CDLL(os.path.join(ROOT, "openairinterface5g_b200", "libldpc_b200.so"))
    info = np.zeros(8, np.int32)
    zs = [z for z in (2, 3, 5, 7, 9, 11, 13, 15) for z in [z * (1 << j) for j in range(8)] if z <= 384 and z % 4 == 0]
    for BG, rates in ((1, (13, 23, 89)), (2, (15, 13, 23))):
        for R in rates:
            for Z in sorted(set(zs)):
                for t in (768, 256):
                    assert lib.nrb200_ldpc_packed_schedule_info(BG, Z, R, t, info.ctypes.data_as(ctypes.c_void_p)) == 0, (BG, Z, R, t)
                    assert info[7] == 1 and info[0] <= t and info[0] % 32 == 0, (BG, Z, R, t, info)
                    assert info[2] == (1 if (Z // 4) % 32 == 0 else 0)
    assert lib.nrb200_ldpc_packed_schedule_info(1, 6, 13, 768, info.ctypes.data_as(ctypes.c_void_p)) == -4       # generic kernel's territory
    for BG, Z, R, t in ((1, 384, 13, 768), (1, 384, 13, 960), (1, 384, 23, 768)):
        lib.nrb200_ldpc_packed_schedule_info(BG, Z, R, t, info.ctypes.data_as(ctypes.c_void_p))
        assert info[3] <= 1.06 * info[4] and info[5] <= 1.06 * info[6], (BG, Z, R, t, info)


def test_ctypes_mirrors_match_the_c_header(tmp_path):
    """sizeof / offsetof of every descriptor in include/nrb200_ldpc.h, as gcc lays them out, against the ctypes mirrors the tests and bench.py call through."""
    import subprocess
    from openairinterface5g_b200 import ldpc
    pairs = [("nrb200_ldpc_dec_params_t", ldpc.DecParams, None), ("nrb200_ldpc_enc_params_t", ldpc.EncParams, None), ("nrb200_decode_abort_t", ldpc.DecodeAbort, None),
             ("nrb200_ldpc_batch_desc_t", ldpc.BatchDesc, "out_stride"), ("nrb200_rm_desc_t", ldpc.RmDesc, None), ("nrb200_pusch_rx_t", ldpc.PuschRxDesc, "d_ptrs_state"),
             ("nrb200_pusch_chest_t", ldpc.PuschChestDesc, "lowpapr_seq"), ("nrb200_pdsch_tx_t", ldpc.PdschTxDesc, "ptrs_re_offset"),
             ("nrb200_rfsim_chan_t", ldpc.RfsimChan, "reserved"), ("nrb200_prach_t", ldpc.PrachDesc, "reserved")]
    ldpc._late_fields()
    pairs += [("nrb200_sch_rx_slot_t", ldpc.SchRxSlotDesc, "seg_payload_bytes"), ("nrb200_sch_rx_bufs_t", ldpc.SchRxBufs, "hard_stride"),
              ("nrb200_pdsch_tx_slot_t", ldpc.PdschTxSlotDesc, "K"), ("nrb200_pdsch_tx_bufs_t", ldpc.PdschTxBufs, "cw_stride")]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "nrb200_slot.h"', '#include "nrb200_rfsim.h"', '#include "nrb200_prach.h"', 'int main(void) {']
    for cname, _, field in pairs:
        src.append(f'  printf("{cname} %zu %zu\\n", sizeof({cname}), {"offsetof(" + cname + ", " + field + ")" if field else "(size_t)0"});')
    src += ['  return 0;', '}']
    c = tmp_path / "abi_layout.c"
    c.write_text("\n".join(src))
    exe = str(tmp_path / "abi_layout")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(c)])
    out = subprocess.check_output([exe], text=True).split("\n")
    got = {l.split()[0]: (int(l.split()[1]), int(l.split()[2])) for l in out if l.strip()}
    for cname, cls, field in pairs:
        assert got[cname][0] == ctypes.sizeof(cls), (cname, got[cname][0], ctypes.sizeof(cls))
        if field:
            assert got[cname][1] == getattr(cls, field).offset, (cname, field, got[cname][1], getattr(cls, field).offset)


def test_sticky_device_rule_matches_the_python_tools():
    """nrb200_sticky_device (the rule the offload convention's C entry points use to pin (ulsch_id, segment) to a GPU) == shard.sticky_gpu (what the
    one-process-per-GPU tools use); host arithmetic only.  Every device index is in range and all devices get work."""
    import ctypes as C
    from openairinterface5g_b200.ldpc import LdpcLib
    from openairinterface5g_b200.shard import sticky_gpu
    lib = LdpcLib().lib
    lib.nrb200_sticky_device.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    for world in (1, 2, 3, 4, 8):
        seen = set()
        for u in range(64):
            for r in range(40):
                d = lib.nrb200_sticky_device(u, r, world)
                assert d == sticky_gpu(u, r, world) and 0 <= d < world
                seen.add(d)
        assert len(seen) == world


def test_pdsch_ptrs_layout_host_logic(oracle):
    """nrb200_pdsch_ptrs_layout / nrb200_pusch_num_llr are host arithmetic (no GPU): PT-RS symbol mask = set_ptrs_symb_idx, PT-RS REs per symbol and the slot's
    LLR count as the pinned oracle leaves them."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from common import PTRS_CASES
    from oracle.bindings import PuschParms, PtrsParms
    from openairinterface5g_b200.ldpc import LdpcLib, PuschRxDesc
    lib = LdpcLib()
    for case in PTRS_CASES:
        N, nb_rx, rb_start, rb_size, Qm, dpos, dtype_, cdm, carrier, start, nsym, L, K, reoff, rnti, slot, nscid, nid = case
        d = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, start, nsym, dpos, dtype_, cdm, 0xFFFFFFFF, 0, 0, 0, rnti, 501, 1, 0, 0, 1)
        d.set_ptrs(L, K, reoff, slot, nscid, nid)
        mask, n = lib.pdsch_ptrs_layout(d)
        assert mask == oracle.ptrs_symbols(start, nsym, L, dpos), case
        z = np.zeros((nb_rx, 14, N, 2), np.int16)
        llr, _, _, nre = oracle.pdsch_rx_slot_ptrs(PuschParms(N, nb_rx, rb_start, 0, rb_size, N - carrier * 6, Qm, dpos, dtype_, cdm), PtrsParms(1, L, K, reoff, rnti, slot, nscid, nid),
                                                   start, nsym, z, z)
        assert [n if (mask >> s) & 1 else 0 for s in range(14)] == nre.tolist() and lib.pusch_num_llr(d) == llr.size, case
