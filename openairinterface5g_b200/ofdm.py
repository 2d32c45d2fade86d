"""Host-side mirror of the frame-parameter arithmetic the slot-level OFDM front end needs (the fields of NR_DL_FRAME_PARMS set by
nr_init_frame_parms, openair1/PHY/INIT/nr_parms.c:160-260, and the rotation tables of nr_modulation.c:586-660), plus the descriptor
builder for nrb200_ofdm_{mod,demod}_slot_* (include/nrb200_dfts.h Part 3).  In an OAI integration these numbers come from `fp`;
here they feed the tests and the benchmark."""
import ctypes as C
import math
import numpy as np


class OfdmSlotDesc(C.Structure):        # nrb200_ofdm_slot_t
    _fields_ = [("fft_size", C.c_uint32), ("n_symb", C.c_uint32), ("n_ant", C.c_uint32), ("f_stride", C.c_uint32), ("t_stride", C.c_uint32),
                ("t_off", C.c_uint32 * 14), ("prefix", C.c_uint32 * 14), ("t_ring", C.c_uint32), ("rotate", C.c_uint32), ("nb_rb", C.c_uint32),
                ("first_carrier_offset", C.c_uint32), ("rot", (C.c_int16 * 2) * 14)]


class NrOfdmParms:
    """ofdm_symbol_size N, numerology mu, N_RB; everything else as nr_init_frame_parms derives it."""

    def __init__(self, N, mu, nb_rb, ofdm_offset_divisor=8):
        self.N, self.mu, self.nb_rb, self.divisor = N, mu, nb_rb, ofdm_offset_divisor
        self.slots_per_subframe = 1 << mu
        self.first_carrier_offset = N - nb_rb * 6
        self.nb_prefix_samples = N // 128 * 9
        self.nb_prefix_samples0 = N // 128 * (9 + (1 << mu))
        p, p0 = self.nb_prefix_samples, self.nb_prefix_samples0
        self.samples_per_slotN0 = (p + N) * 14
        self.samples_per_slot0 = p0 + 13 * p + 14 * N
        self.samples_per_subframe = (p0 + N) * 2 + (p + N) * (14 * self.slots_per_subframe - 2)
        self.samples_per_frame = 10 * self.samples_per_subframe

    def samples_per_slot(self, slot):
        if self.mu == 0:
            return self.samples_per_subframe
        return self.samples_per_slotN0 if slot % (self.slots_per_subframe // 2) else self.samples_per_slot0

    def slot_timestamp(self, slot):
        return sum(self.samples_per_slot(s) for s in range(slot))

    def slot_geometry(self, slot):
        """(prefix[14], cp_start[14]) relative to the slot start: the longer CP goes to every (7 << mu)-th symbol of the subframe."""
        prefix, start, pos = [], [], 0
        for l in range(14):
            cp = self.nb_prefix_samples if (slot * 14 + l) % (7 << self.mu) else self.nb_prefix_samples0
            prefix.append(cp); start.append(pos)
            pos += cp + self.N
        return prefix, start

    def symbol_rotation(self, f0):
        """init_symbol_rotation (nr_modulation.c:586-637): 14 << mu entries {re, im}."""
        f32 = np.float32
        inv = float(f32(1) / f32(1 << self.mu))
        Tc = (1 / 480e3 / 4096)
        Nu = 2048 * 64 * inv
        Ncp0 = 16 * 64 + (144 * 64 * inv)
        Ncp1 = (144 * 64 * inv)
        out = np.zeros((14 << self.mu, 2), np.int16)
        tl = 0.0
        for l in range(14 << self.mu):
            Ncp = Ncp0 if (l == 0 or l == 7 * (1 << self.mu)) else Ncp1
            poff = 2 * math.pi * (tl + (Ncp * Tc)) * f0
            out[l, 0] = math.floor(math.cos(poff) * 32767)
            out[l, 1] = math.floor(math.sin(-poff) * 32767)
            tl += (Nu + Ncp) * Tc
        return out

    def timeshift_rotation(self):
        """init_timeshift_rotation (nr_modulation.c:639-660)."""
        so = self.nb_prefix_samples // self.divisor
        out = np.zeros((self.N, 2), np.int16)
        for i in range(self.N):
            poff = -i * 2.0 * math.pi * so / self.N
            out[i, 0] = _c_round(math.cos(poff) * 32767)
            out[i, 1] = _c_round(math.sin(-poff) * 32767)
        return out

    def rx_window_offsets(self, slot, sample_offset=0):
        """First sample of each symbol's FFT window in the frame ring (nr_slot_fep_ul, slot_fep_nr.c:238-247), modulo the frame."""
        prefix, start = self.slot_geometry(slot)
        ss = self.slot_timestamp(slot)
        back = self.nb_prefix_samples // self.divisor
        return [(ss + start[l] + prefix[l] - back - sample_offset) % self.samples_per_frame for l in range(14)]

    def desc(self, slot, n_ant, rot=None, rx=False, sample_offset=0, f_stride=None, t_stride=None, t_base=0):
        """Descriptor for the 14 symbols of `slot`.  TX: t_off relative to the slot start + t_base; RX: ring offsets."""
        d = OfdmSlotDesc()
        d.fft_size, d.n_symb, d.n_ant = self.N, 14, n_ant
        d.nb_rb, d.first_carrier_offset = self.nb_rb, self.first_carrier_offset
        prefix, start = self.slot_geometry(slot)
        offs = self.rx_window_offsets(slot, sample_offset) if rx else [t_base + s for s in start]
        for l in range(14):
            d.t_off[l], d.prefix[l] = offs[l], prefix[l]
        d.t_ring = self.samples_per_frame if rx else 0
        d.f_stride = 14 * self.N if f_stride is None else f_stride
        d.t_stride = (self.samples_per_frame if rx else start[13] + prefix[13] + self.N) if t_stride is None else t_stride
        d.rotate = 0 if rot is None else 1
        if rot is not None:
            so = (slot % self.slots_per_subframe) * 14
            for l in range(14):
                d.rot[l][0], d.rot[l][1] = int(rot[so + l][0]), int(rot[so + l][1])
        return d


def _c_round(x):
    """C round(): halves away from zero."""
    return int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)
