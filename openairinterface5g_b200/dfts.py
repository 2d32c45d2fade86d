"""Host-side binding of libdfts_b200.so -- the B200-native drop-in for OAI's loadable DFT library (`dft`/`idft` function-pointer
pair, openair1/PHY/TOOLS/dfts_load.c:47-61, tools_defs.h:404-676).  Bit-exact Q15 arithmetic of oai_dfts.c for the OFDM sizes."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdfts_b200.so")

DFT_SIZES = [12, 24, 36, 48, 60, 64, 72, 96, 108, 120, 128, 144, 180, 192, 216, 240, 256, 288, 300, 324, 360, 384, 432, 480, 512, 540, 576, 600, 648, 720,
             768, 864, 900, 960, 972, 1024, 1080, 1152, 1200, 1296, 1440, 1500, 1536, 1620, 1728, 1800, 1920, 1944, 2048, 2160, 2304, 2400, 2592, 2700,
             2880, 2916, 3000, 3072, 3240, 4096, 6144, 8192, 9216, 12288, 18432, 24576, 36864, 49152, 73728, 98304]      # FOREACH_DFTSZ
IDFT_SIZES = [64, 128, 256, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192, 9216, 12288, 16384, 18432, 24576, 32768, 36864, 49152, 65536, 73728,
              98304]                                                                                                      # FOREACH_IDFTSZ
SUPPORTED = [64, 128, 256, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192]          # one transform per call, both directions
# the DFT-s-OFDM family (PUSCH transform precoding): forward only, every call transforms FOUR interleaved sequences (c16 number 4 n + l = element n of
# transform l, oai_dfts.c:4352), i.e. 4 N c16 in and out
FOURWAY = [N for N in DFT_SIZES if N <= 3240 and N not in SUPPORTED]
LARGE = [12288, 16384, 18432, 24576, 32768, 36864, 49152, 65536, 98304]               # 65536: inverse only; 9216 and 73728 do not exist in the reference either


def c16_per_call(N):
    return 4 * N if N in FOURWAY else N


def get_dft(N):
    """dft_size_idx_t of N (tools_defs.h:547-610)."""
    return DFT_SIZES.index(N)


def get_idft(N):
    return IDFT_SIZES.index(N)


class DftsLib:
    def __init__(self, path=_SO):
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = self.lib = C.CDLL(path)
        L.dft.argtypes = [C.c_uint8, C.c_void_p, C.c_void_p, C.c_ubyte]
        L.dft.restype = None
        L.idft.argtypes = [C.c_uint8, C.c_void_p, C.c_void_p, C.c_ubyte]
        L.idft.restype = None
        L.nrb200_dft_batch_dev.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.nrb200_dft_batch_host.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.nrb200_ofdm_mod_slot_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_ofdm_demod_slot_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_ofdm_mod_slot_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_ofdm_demod_slot_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrb200_dfts_last_error.restype = C.c_char_p
        L.nrb200_dfts_launch_count.restype = C.c_uint64

    def autoinit(self):
        if self.lib.dfts_autoinit() != 0:
            raise RuntimeError("dfts_autoinit failed: " + (self.lib.nrb200_dfts_last_error() or b"").decode())

    def dft(self, sizeidx, x, scale=1):
        """The plug-in call OAI makes: dft(get_dft(N), in, out, scale_flag), one transform."""
        x = np.ascontiguousarray(x, dtype=np.int16)
        y = np.zeros_like(x)
        self.lib.dft(sizeidx, x.ctypes.data, y.ctypes.data, scale)
        return y

    def idft(self, sizeidx, x, scale=1):
        x = np.ascontiguousarray(x, dtype=np.int16)
        y = np.zeros_like(x)
        self.lib.idft(sizeidx, x.ctypes.data, y.ctypes.data, scale)
        return y

    def batch_host(self, N, inverse, x, scale=1):
        x = np.ascontiguousarray(x, dtype=np.int16)
        n = x.size // (2 * c16_per_call(N))
        assert n * 2 * c16_per_call(N) == x.size
        y = np.zeros_like(x)
        rc = self.lib.nrb200_dft_batch_host(N, int(inverse), n, x.ctypes.data, y.ctypes.data, scale)
        if rc != 0:
            raise RuntimeError(f"nrb200_dft_batch_host rc={rc}: " + (self.lib.nrb200_dfts_last_error() or b"").decode())
        return y

    def batch_torch(self, N, inverse, x, scale=1, out=None):
        import torch
        assert x.is_cuda and x.dtype == torch.int16 and x.is_contiguous()
        n = x.numel() // (2 * c16_per_call(N))
        if out is None:
            out = torch.empty_like(x)
        rc = self.lib.nrb200_dft_batch_dev(N, int(inverse), n, x.data_ptr(), out.data_ptr(), scale, torch.cuda.current_stream(x.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"nrb200_dft_batch_dev rc={rc}")
        return out

    # ---- slot-level OFDM front end (include/nrb200_dfts.h Part 3)
    def _err(self, what, rc):
        raise RuntimeError(f"{what} rc={rc}: " + (self.lib.nrb200_dfts_last_error() or b"").decode())

    def ofdm_mod_slot_host(self, parms, slot, txdataF, rot=None):
        """txdataF[n_ant][14*N*2] int16 -> txdata[n_ant][slot samples * 2]; apply_nr_rotation_TX + PHY_ofdm_mod in one launch."""
        F = np.ascontiguousarray(txdataF, dtype=np.int16).reshape(len(txdataF), -1)
        na = F.shape[0]
        d = parms.desc(slot, na, rot)
        out = np.zeros((na, 2 * d.t_stride), np.int16)
        pin = (C.c_void_p * na)(*[F[a].ctypes.data for a in range(na)])
        pout = (C.c_void_p * na)(*[out[a].ctypes.data for a in range(na)])
        rc = self.lib.nrb200_ofdm_mod_slot_host(C.addressof(d), pin, pout)
        if rc != 0:
            self._err("nrb200_ofdm_mod_slot_host", rc)
        return out

    def ofdm_demod_slot_host(self, parms, slot, rxdata, rot=None, sample_offset=0):
        """rxdata[n_ant][samples_per_frame*2] (frame ring) -> rxdataF[n_ant][14*N*2]; nr_slot_fep_ul x 14 + apply_nr_rotation_RX."""
        X = np.ascontiguousarray(rxdata, dtype=np.int16).reshape(len(rxdata), -1)
        na = X.shape[0]
        d = parms.desc(slot, na, rot, rx=True, sample_offset=sample_offset)
        ts = np.ascontiguousarray(parms.timeshift_rotation()) if rot is not None else None
        out = np.zeros((na, 2 * 14 * parms.N), np.int16)
        pin = (C.c_void_p * na)(*[X[a].ctypes.data for a in range(na)])
        pout = (C.c_void_p * na)(*[out[a].ctypes.data for a in range(na)])
        rc = self.lib.nrb200_ofdm_demod_slot_host(C.addressof(d), pin, None if ts is None else ts.ctypes.data, pout)
        if rc != 0:
            self._err("nrb200_ofdm_demod_slot_host", rc)
        return out

    def ofdm_mod_slot_torch(self, desc, txdataF, txdata):
        import torch
        rc = self.lib.nrb200_ofdm_mod_slot_dev(C.addressof(desc), txdataF.data_ptr(), txdata.data_ptr(), torch.cuda.current_stream(txdataF.device).cuda_stream)
        if rc != 0:
            self._err("nrb200_ofdm_mod_slot_dev", rc)
        return txdata

    def ofdm_demod_slot_torch(self, desc, rxdata, timeshift, rxdataF):
        import torch
        rc = self.lib.nrb200_ofdm_demod_slot_dev(C.addressof(desc), rxdata.data_ptr(), 0 if timeshift is None else timeshift.data_ptr(), rxdataF.data_ptr(),
                                                 torch.cuda.current_stream(rxdata.device).cuda_stream)
        if rc != 0:
            self._err("nrb200_ofdm_demod_slot_dev", rc)
        return rxdataF

    def launch_count(self):
        return int(self.lib.nrb200_dfts_launch_count())


_lib = None


def load_dftslib():
    """load_dftslib() equivalent (dfts_load.c:47-61): dlopen + dfts_autoinit."""
    global _lib
    if _lib is None:
        _lib = DftsLib()
        _lib.autoinit()
    return _lib
