"""Host-side transport-block arithmetic around the codec -- the scalar bookkeeping the reference does on the CPU before it calls the
coding library (it stays on the host here too; the per-bit work is in the CUDA kernels):
  nr_segmentation           openair1/PHY/CODING/nr_segmentation.c:32-180
  nr_get_G / nr_get_E       openair1/PHY/NR_TRANSPORT/nr_tbs_tools.c:37-62
  nr_get_R_ldpc_decoder     openair1/PHY/CODING/nr_rate_matching.c:390-422
tests/test_golden_oracle.py checks them against the oracle restatement (pinned to the compiled reference)."""
import numpy as np

_STEPS = ((16, 1), (32, 2), (64, 4), (128, 8), (256, 16), (384, 32))
K0_NUM = {1: (0, 17, 33, 56), 2: (0, 13, 25, 43)}          # nr_rate_matching.c:34


def nr_segmentation(B, BG):
    """B = transport block size in bits including the TB CRC.  Returns dict(C, K, Z, F, Kprime, L)."""
    Kcb = 8448 if BG == 1 else 3840
    if B <= Kcb:
        L, Cn, Bp = 0, 1, B
    else:
        L = 24
        Cn = B // (Kcb - L)
        if (Kcb - L) * Cn < B:
            Cn += 1
        Bp = B + Cn * L
    Kp = Bp // Cn
    Kb = 22 if BG == 1 else (10 if B > 640 else 9 if B > 560 else 8 if B > 192 else 6)
    Zmin = Kp // Kb + (1 if Kp % Kb else 0)
    if Zmin <= 2:
        Z = 2
    elif Zmin <= 16:
        Z = Zmin
    else:
        Z = None
        for hi, step in _STEPS[1:]:
            if Zmin <= hi:
                Z = (Zmin // step) * step
                if Z < Zmin:
                    Z += step
                break
        if Z is None:
            raise ValueError("transport block too large for one code block set")
    K = Z * (22 if BG == 1 else 10)
    return {"C": Cn, "K": K, "Z": Z, "F": K - Kp, "Kprime": Kp, "L": L, "Kb": Kb}


def nr_get_G(nb_rb, nb_symb_sch, nb_re_dmrs, length_dmrs, unav_res, Qm, Nl):
    return ((12 * nb_symb_sch) - (nb_re_dmrs * length_dmrs)) * nb_rb * Qm * Nl - unav_res * Qm * Nl


def nr_get_E(G, C, Qm, Nl, r):
    if r <= C - ((G // (Nl * Qm)) % C) - 1:
        return Nl * Qm * (G // (Nl * Qm * C))
    return Nl * Qm * ((G // (Nl * Qm * C)) + 1)


def nr_get_R_ldpc_decoder(rv, E, BG, Z, llrLen=0, round_=0):
    """Returns (R, llrLen): the decoder's rate LUT selector and the running HARQ length."""
    Ncb = (66 if BG == 1 else 50) * Z
    info = K0_NUM[BG][rv] * Z + E
    if round_ == 0:
        llrLen = info
    info = min(info, Ncb)
    llrLen = max(llrLen, info)
    sys_bits = (22 if BG == 1 else 10) * Z
    R = float(np.float32(sys_bits) / np.float32(info + 2 * Z))
    if BG == 2:
        return (15 if R < 0.3333 else 13 if R < 0.6667 else 23), llrLen
    return (13 if R < 0.6667 else 23 if R < 0.8889 else 89), llrLen


def segment_transport_block(lib, tb_with_crc, BG):
    """Byte-level nr_segmentation: tb_with_crc = payload followed by its CRC (uint8, MSB-first bits).  Returns (segments uint8 (C, K/8),
    seg) with per-segment CRC24B attached when C > 1 and zero filler bytes.  `lib` computes the CRCs on the device (crc_batch_host)."""
    tb = np.ascontiguousarray(tb_with_crc, dtype=np.uint8)
    seg = nr_segmentation(tb.size * 8, BG)
    Cn, K, Kp, L = seg["C"], seg["K"], seg["Kprime"], seg["L"]
    out = np.zeros((Cn, K // 8), dtype=np.uint8)
    nbytes = (Kp - L) >> 3
    out[:, :nbytes] = tb[:Cn * nbytes].reshape(Cn, nbytes)
    if Cn > 1:
        crc = lib.crc_batch_host(1, np.ascontiguousarray(out[:, :nbytes]), Kp - L) >> 8            # crc24b, returned left aligned
        out[:, nbytes] = (crc >> 16) & 0xFF
        out[:, nbytes + 1] = (crc >> 8) & 0xFF
        out[:, nbytes + 2] = crc & 0xFF
    return out, seg
