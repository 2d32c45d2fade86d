"""Device-resident PDSCH slot chain (the nr_dlsim path, BASELINE config 3: 100 MHz, 273 PRB, 2 x 2, two layers, MCS 28) built from the library's kernels.
The order of calls mirrors the reference's procedures:
  gNB transmit : nr_generate_pdsch (NR_TRANSPORT/nr_dlsch.c:56-583) = nr_dlsch_encoding (TB CRC24A, nr_segmentation + CRC24B, LDPC encode, rate matching +
                 interleaving; nr_dlsch_coding.c:280-420) -> scrambling -> modulation -> layer mapping -> DMRS + resource mapping -> precoding, then
                 nr_feptx0 (apply_nr_rotation_TX + PHY_ofdm_mod, SCHED_NR/nr_ru_procedures.c:55-140)
  UE receive   : nr_slot_fep (MODULATION/slot_fep_nr.c:37-113) -> nr_pdsch_channel_estimation per DMRS port (NR_UE_ESTIMATION/nr_dl_channel_estimation.c:1614)
                 -> nr_rx_pdsch (level, compensation, MRC, zero forcing, LLRs, layer de-mapping; NR_UE_TRANSPORT/nr_dlsch_demodulation.c:241-684)
                 -> nr_dlsch_unscrambling -> nr_dlsch_decoding (de-interleave / rate recover -> LDPC decode with CRC24B stop -> TB CRC;
                 NR_UE_TRANSPORT/nr_dlsch_decoding.c:235-420), as nr_ue_pdsch_procedures / nr_ue_dlsch_procedures sequence them
                 (SCHED_NR_UE/phy_procedures_nr_ue.c:520-760).
Between the two sits nr_dlsim's double-precision channel model (multipath_channel + add_noise, dlsim.c:1098-1099), which is simulator code and not on the path:
`channel()` is a small flat 2 x 2 mix plus noise in torch and is never timed.  torch is used for buffers, the byte plumbing of the segmentation and the channel."""
import numpy as np
import torch

from . import transport as T
from . import ldpc as _L
from .ldpc import CRC24_B, PdschTxDesc, PuschChestDesc, PuschRxDesc
from .ofdm import NrOfdmParms


class PdschSlotChain:
    def __init__(self, lib, dl, device, A=434280, N=4096, mu=1, carrier_rb=273, rb_start=0, rb_size=273, nb_ant=2, Qm=6, slot=1, rnti=0x1234, nid=77,
                 dl_freq=3619200000.0, max_iter=8, dmrs_id=55, n_layers=2, tx_amp=512, start_symbol=1, nr_symbols=13, ptrs=None):
        """ptrs = (PTRSTimeDensity (log2 of L), PTRSFreqDensity K, PTRSReOffset) switches PT-RS on at both ends (pduBitmap & 1; one layer: the UE side of the
        reference handles layer 0 only): the gNB inserts the pilots, the encoder's G shrinks by unav_res, the UE estimates / interpolates / compensates the common
        phase error inside its slot receiver."""
        self.lib, self.dl, self.dev = lib, dl, device
        self.latency_mode = 1            # decoder: a cluster of SMs per code block (one slot alone); the pipelines below run many slots and set 0
        self.P = NrOfdmParms(N, mu, carrier_rb)
        self.N, self.nb, self.Qm, self.slot, self.rnti, self.nid, self.max_iter, self.nl = N, nb_ant, Qm, slot, rnti, nid, max_iter, n_layers
        self.rb_start, self.rb_size, self.A = rb_start, rb_size, A
        assert n_layers in (1, 2, 3, 4) and nb_ant >= n_layers
        self.dmrs_pos, self.dmrs_type, self.cdm = 1 << 2, 0, 2                     # one type-1 DMRS symbol (ports 0, 1 share CDM group 0), no data on it
        self.seg = T.nr_segmentation(A + 24, 1)
        assert (A + 24 + self.seg["C"] * self.seg["L"]) % (8 * self.seg["C"]) == 0, "pick A like a real TBS: whole bytes per segment"
        self.C, self.K, self.Z, self.F = self.seg["C"], self.seg["K"], self.seg["Z"], self.seg["F"]
        fco = self.P.first_carrier_offset
        self.txd = PdschTxDesc(N, nb_ant, slot, rb_start, 0, rb_size, fco, Qm, n_layers, start_symbol, nr_symbols, self.dmrs_pos, self.dmrs_type, self.cdm,
                               (1 << n_layers) - 1, 0, dmrs_id, nid, rnti, tx_amp, 14 * N)
        self.ptrs = ptrs
        unav_res = 0
        if ptrs is not None:
            assert n_layers == 1
            self.txd.set_ptrs(*ptrs)
        self.G = lib.pdsch_tx_num_bits(self.txd)
        if ptrs is not None:
            unav_res = (T.nr_get_G(rb_size, nr_symbols, 12, 1, 0, Qm, n_layers) - self.G) // (Qm * n_layers)      # harq->unav_res (nr_dlsch.c:111)
        assert self.G == T.nr_get_G(rb_size, nr_symbols, 12, 1, unav_res, Qm, n_layers) and self.G > 0
        E = [T.nr_get_E(self.G, self.C, Qm, n_layers, r) for r in range(self.C)]
        self.R = T.nr_get_R_ldpc_decoder(0, E[0], 1, self.Z)[0]
        self.E = torch.tensor(E, dtype=torch.int32, device=device)
        self.Eoff = torch.tensor(np.concatenate([[0], np.cumsum(E)[:-1]]), dtype=torch.int32, device=device)
        self.rot = self.P.symbol_rotation(dl_freq)
        self.ts = torch.from_numpy(self.P.timeshift_rotation()).to(device)
        # ---- transmit-side buffers
        self.nbytes = (self.seg["Kprime"] - self.seg["L"]) // 8                      # payload bytes per segment
        self.tb_tx = torch.zeros((1, (A + 24) // 8), dtype=torch.uint8, device=device)
        self.crc1 = torch.empty(1, dtype=torch.int32, device=device)
        self.crcC = torch.empty(self.C, dtype=torch.int32, device=device)
        self.segs = torch.zeros((self.C, self.K // 8), dtype=torch.uint8, device=device)
        self.cw = torch.empty((self.C, 66 * self.Z), dtype=torch.uint8, device=device)
        self.f = torch.empty(self.G, dtype=torch.uint8, device=device)
        self.txF = torch.zeros((nb_ant, 14 * N, 2), dtype=torch.int16, device=device)
        self.dtx = self.P.desc(slot, nb_ant, self.rot)
        self.txdata = torch.zeros((nb_ant, self.dtx.t_stride, 2), dtype=torch.int16, device=device)
        self.shifts = torch.tensor([24, 16, 8], dtype=torch.int32, device=device)
        # ---- receive-side buffers
        self.rxd = PuschRxDesc(N, nb_ant, rb_start, 0, rb_size, fco, Qm, start_symbol, nr_symbols, self.dmrs_pos, self.dmrs_type, self.cdm,
                               0, 14 * N, 14 * N, 1, rnti, nid, n_layers, 0, 0, 1)
        if ptrs is not None:
            self.ptrs_state = torch.zeros(32, dtype=torch.int32, device=device)     # ptrs_phase_per_slot[0] as the library leaves it, + status
            self.rxd.set_ptrs(ptrs[0], ptrs[1], ptrs[2], slot, 0, dmrs_id, self.ptrs_state.data_ptr())
            mask, n_re = lib.pdsch_ptrs_layout(self.rxd)
            assert bin(mask).count("1") * n_re == unav_res
        assert lib.pusch_num_llr(self.rxd) == self.G
        self.cdesc = PuschChestDesc(N, nb_ant, slot, 2, 0, rb_start, 0, rb_size, fco, 0, dmrs_id, 14 * N, 14 * N, min(n_layers, 2), 1)   # UE estimator, ports 0 (and 1) in one call
        # layers 3 and 4: ports 2 (and 3) of the second CDM group, a second call
        self.cdesc2 = PuschChestDesc(N, nb_ant, slot, 2, 2, rb_start, 0, rb_size, fco, 0, dmrs_id, 14 * N, 14 * N, n_layers - 2, 1) if n_layers > 2 else None
        self.est = torch.zeros((n_layers * nb_ant, 14 * N, 2), dtype=torch.int16, device=device)      # dl_ch_estimates[p * nb_rx + aarx]
        self.chest_scratch = torch.empty(lib.pusch_chest_scratch_bytes(self.cdesc), dtype=torch.uint8, device=device)
        self.chest_state = torch.zeros((n_layers, 18), dtype=torch.int32, device=device)
        self.rxF = torch.empty((nb_ant, 14 * N, 2), dtype=torch.int16, device=device)
        self.level = torch.zeros(9, dtype=torch.int32, device=device)
        self.llr16 = torch.empty(self.G, dtype=torch.int16, device=device)
        self.harq = torch.zeros((self.C, 66 * self.Z), dtype=torch.int16, device=device)
        self.llr8 = torch.empty((self.C, 68 * self.Z), dtype=torch.int8, device=device)
        self.hard = torch.empty((self.C, 68 * self.Z // 8), dtype=torch.uint8, device=device)
        self.iters = torch.empty(self.C, dtype=torch.int32, device=device)
        self.tb = torch.empty((1, (A + 24) // 8), dtype=torch.uint8, device=device)
        self.tbcrc = torch.empty(1, dtype=torch.int32, device=device)
        self.drx = self.P.desc(slot, nb_ant, self.rot, rx=True)

    # ------------------------------------------------------------------ gNB transmit chain (timed)
    def _c_tx(self, payload):
        _L._late_fields()
        C_ = _L.C
        d = _L.PdschTxSlotDesc()
        C_.memmove(C_.addressof(d.tx), C_.addressof(self.txd), C_.sizeof(self.txd))
        C_.memmove(C_.addressof(d.ofdm), C_.addressof(self.dtx), C_.sizeof(self.dtx))
        d.rm = self.lib._rmdesc(1, self.Z, self.Qm, 0, self.C, 0, self.F, self.C)
        d.A, d.K = self.A, self.K
        b = _L.PdschTxBufs()
        b.d_payload, b.d_segs, b.d_seg_scratch, b.d_cw = payload.data_ptr(), self.segs.data_ptr(), self.crc1.data_ptr(), self.cw.data_ptr()
        b.d_E, b.d_Eoff, b.d_f, b.d_txdataF, b.d_txdata = self.E.data_ptr(), self.Eoff.data_ptr(), self.f.data_ptr(), self.txF.data_ptr(), self.txdata.data_ptr()
        b.seg_stride, b.cw_stride = self.segs.shape[1], self.cw.shape[1]
        return d, b

    def _c_rx(self, rxdata):
        _L._late_fields()
        C_ = _L.C
        d = _L.SchRxSlotDesc()
        C_.memmove(C_.addressof(d.ofdm), C_.addressof(self.drx), C_.sizeof(self.drx))
        C_.memmove(C_.addressof(d.chest), C_.addressof(self.cdesc), C_.sizeof(self.cdesc))
        C_.memmove(C_.addressof(d.rx), C_.addressof(self.rxd), C_.sizeof(self.rxd))
        d.rm = self.lib._rmdesc(1, self.Z, self.Qm, 0, self.C, 0, self.F, self.C, 1)
        d.R, d.numMaxIter, d.use_estimates, d.latency_mode = self.R, self.max_iter, 0, self.latency_mode
        d.crc_len_bits, d.seg_crc_type = self.K - self.F, CRC24_B
        d.A, d.tb_crc_bits, d.seg_payload_bytes = self.A, 24, self.nbytes
        b = _L.SchRxBufs()
        b.d_rxdata, b.d_timeshift, b.d_rxdataF, b.d_est = rxdata.data_ptr(), self.ts.data_ptr(), self.rxF.data_ptr(), self.est.data_ptr()
        b.d_chest_scratch, b.d_chest_state, b.d_level, b.d_llr16 = self.chest_scratch.data_ptr(), self.chest_state.data_ptr(), self.level.data_ptr(), self.llr16.data_ptr()
        b.d_E, b.d_Eoff, b.d_harq, b.d_llr8, b.d_hard = self.E.data_ptr(), self.Eoff.data_ptr(), self.harq.data_ptr(), self.llr8.data_ptr(), self.hard.data_ptr()
        b.d_iters, b.d_tb, b.d_tbcrc = self.iters.data_ptr(), self.tb.data_ptr(), self.tbcrc.data_ptr()
        b.harq_stride, b.llr8_stride, b.hard_stride = self.harq.shape[1], self.llr8.shape[1], self.hard.shape[1]
        return d, b

    def transmit(self, payload, staged=False):
        """payload: uint8[A / 8] on the device.  Returns the slot's time-domain samples int16 [nb_tx, samples, 2].  ONE library call
        (nrb200_pdsch_slot_tx_dev); staged=True issues the stages one entry point at a time from here (compared in the tests)."""
        lib, dl = self.lib, self.dl
        if not staged:
            d, b = self._c_tx(payload)
            self._keep_tx = (d, b)
            lib.pdsch_slot_tx_torch(d, b, self.dev)
            return self.txdata
        lib.tb_segment_torch(1, self.A, payload, self.segs, self.crc1)                      # TB CRC24A + nr_segmentation + CRC24B per segment
        lib.encode_batch_torch(1, self.Z, self.K, self.segs, out=self.cw)
        lib.rm_tx_torch(1, self.Z, self.Qm, 0, self.C, 0, self.F, self.cw, self.E, self.Eoff, self.f)
        lib.pdsch_tx_slot_torch(self.txd, self.f, self.txF)                                 # scrambling ... txdataF in one launch
        dl.ofdm_mod_slot_torch(self.dtx, self.txF, self.txdata)                             # rotation + IDFT + CP
        return self.txdata

    # ------------------------------------------------------------------ the simulator's channel (never timed)
    def channel(self, txdata, seed=1, snr_db=35.0, gain=3.0, coupling=0.15, cpe_per_symbol=0.0):
        """Flat nb_rx x nb_tx mix + white noise in the time domain, placed at the slot's position of a frame buffer.  Returns int16 [nb_rx, samples_per_frame, 2].
        The default gain puts the receiver's power-of-two LLR scaling (log2_maxh) where ~40 % of the 64QAM LLRs sit at the int8 rail: with the reference's
        unscaled min-sum and the double-amplitude REs of its resource mapper (DESIGN.md defect 9) that is the scaling at which every 273-PRB slot decodes within
        8 iterations (profiles/dlslot_oppoint_r01l.jsonl); a factor sqrt(2) either way and 3-10 % of the code blocks need more."""
        dev, nb = self.dev, self.nb
        g = torch.Generator(device=dev); g.manual_seed(seed)
        x = torch.view_as_complex(txdata.to(torch.float32).contiguous())
        ph = torch.rand((nb, nb), generator=g, device=dev) * 6.2831853
        H = torch.polar(torch.full((nb, nb), coupling, device=dev) + (1.0 - coupling) * torch.eye(nb, device=dev), ph) * gain
        y = H.to(torch.complex64) @ x
        if cpe_per_symbol:                                                  # a slow common phase drift (rad per OFDM symbol): what PT-RS is there to track
            n = torch.arange(y.shape[1], device=dev, dtype=torch.float32) * (cpe_per_symbol / (self.N * 1.0703125))
            y = y * torch.polar(torch.ones_like(n), n)
        sig = torch.sqrt(torch.mean(torch.abs(y) ** 2)) * 10.0 ** (-snr_db / 20.0) * 0.70711
        yr = torch.view_as_real(y) + sig * torch.randn(y.shape + (2,), generator=g, device=dev)
        rxdata = torch.zeros((nb, self.P.samples_per_frame, 2), dtype=torch.int16, device=dev)
        ss = self.P.slot_timestamp(self.slot)
        rxdata[:, ss:ss + txdata.shape[1]] = torch.clamp(torch.round(yr), -32768, 32767).to(torch.int16)
        return rxdata

    # ------------------------------------------------------------------ UE receive chain (timed)
    def receive(self, rxdata, staged=False):
        lib, dl = self.lib, self.dl
        if not staged:
            d, b = self._c_rx(rxdata)
            self._keep_rx = (d, b)
            lib.sch_slot_rx_torch(d, b, self.dev)
            return self.tb, self.iters, self.tbcrc
        dl.ofdm_demod_slot_torch(self.drx, rxdata, self.ts, self.rxF)                       # nr_slot_fep x 14
        lib.pusch_chest_torch(self.cdesc, self.rxF, self.est, self.chest_scratch, self.chest_state)   # nr_pdsch_channel_estimation, every port
        if self.cdesc2 is not None:
            lib.pusch_chest_torch(self.cdesc2, self.rxF, self.est[2 * self.nb:], self.chest_scratch, self.chest_state[2:])
        lib.pusch_inner_rx_torch(self.rxd, self.rxF, self.est, self.llr16, level=self.level)   # nr_rx_pdsch (+ unscrambling)
        lib.rm_rx_torch(1, self.Z, self.Qm, 0, self.C, 0, self.F, self.llr16, self.E, self.Eoff, self.harq, self.llr8, clear=1)
        lib.decode_batch_torch(1, self.Z, self.R, self.max_iter, self.llr8, use_crc=1, crc_len_bits=self.K - self.F, crc_type=CRC24_B,
                               out=self.hard, iters=self.iters, latency_mode=self.latency_mode)
        self.tb.view(-1).copy_(self.hard[:, :self.nbytes].reshape(-1))
        lib.crc_batch_torch(0, self.tb, self.A + 24, out=self.tbcrc)
        return self.tb, self.iters, self.tbcrc


class PdschSlotPipeline:
    """K PDSCH slots in flight on one GPU ("one CUDA stream per transport block"): K independent PdschSlotChain instances (own buffers, own stream), each slot's
    gNB transmit + UE receive launches captured once into a CUDA graph and replayed.  This is how the library is meant to be driven when slots/s rather than the
    latency of one slot is the figure of merit: nr_dlsim's own throughput mode is one process per core, each working through independent slots
    (cmake_targets/autotests run the physims in parallel the same way); a single slot's 52 code blocks occupy a third of the SMs."""

    def __init__(self, lib, dl, device, n_inflight, use_graphs=True, seed0=100, **cfg):
        self.K, self.dev, self.use_graphs = n_inflight, device, use_graphs
        self.chains, self.graphs, self.streams, self.payload, self.rx, self.h_payload, self.h_tb = [], [], [], [], [], [], []
        for k in range(n_inflight):
            s = torch.cuda.Stream(device=device)
            with torch.cuda.stream(s):
                ch = PdschSlotChain(lib, dl, device, **cfg)
                ch.latency_mode = 0                                  # many slots in flight: one CTA per code block spends the fewest SM-cycles
                hp = torch.from_numpy(np.random.default_rng(seed0 + k).integers(0, 256, size=ch.A // 8, dtype=np.uint8)).pin_memory()
                p = hp.to(device)
                rx = ch.channel(ch.transmit(p), seed=seed0 + k)
                ch.receive(rx)                                       # warm-up: lazy attribute setup and table uploads happen outside the capture
                s.synchronize()
                g = None
                if use_graphs:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=s):
                        ch.transmit(p)
                        ch.receive(rx)
            self.chains.append(ch); self.graphs.append(g); self.streams.append(s); self.payload.append(p); self.rx.append(rx); self.h_payload.append(hp)
            self.h_tb.append(torch.empty_like(ch.tb, device="cpu").pin_memory())
        torch.cuda.synchronize(device)

    def round(self, e2e=False):
        """One slot on every stream.  e2e: the payload comes from pinned host memory and the decoded transport block goes back, on the slot's stream."""
        for k in range(self.K):
            with torch.cuda.stream(self.streams[k]):
                if e2e:
                    self.payload[k].copy_(self.h_payload[k], non_blocking=True)
                if self.graphs[k] is not None:
                    self.graphs[k].replay()
                else:
                    self.chains[k].transmit(self.payload[k]); self.chains[k].receive(self.rx[k])
                if e2e:
                    self.h_tb[k].copy_(self.chains[k].tb, non_blocking=True)

    def fork(self, ev):
        for s in self.streams:
            s.wait_event(ev)

    def join(self, cur):
        for s in self.streams:
            ev = torch.cuda.Event(); ev.record(s); cur.wait_event(ev)

    def timed_rounds(self, n_rounds, e2e=False, warm=3):
        """CUDA-event time of n_rounds x K slots (events on the current stream, which every slot stream forks from and joins back into).  Returns ms."""
        for _ in range(warm):
            self.round(e2e)
        torch.cuda.synchronize(self.dev)
        cur = torch.cuda.current_stream(self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        self.fork(e0)
        for _ in range(n_rounds):
            self.round(e2e)
        self.join(cur)
        e1.record(cur)
        torch.cuda.synchronize(self.dev)
        return e0.elapsed_time(e1)

    def check(self, host=False):
        """Every slot decoded: all code blocks within the iteration cap, TB CRC zero, transport block equal to the payload (host: as read back by the e2e copies)."""
        ok = []
        for k, ch in enumerate(self.chains):
            tb = self.h_tb[k] if host else ch.tb.cpu()
            ok.append(bool((ch.iters <= ch.max_iter).all()) and int(ch.tbcrc.cpu()[0]) == 0 and bool((tb.view(-1)[:self.h_payload[k].numel()] == self.h_payload[k]).all()))
        return ok
