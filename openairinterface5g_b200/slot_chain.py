"""Device-resident PUSCH slot chains built from the library's kernels -- the order of calls mirrors the reference's procedures:
  receive : nr_fep_full (OFDM demodulation) -> nr_rx_pusch_tp (level, compensation, LLR, descrambling) -> nr_ulsch_decoding
            (de-interleave / rate recover / HARQ combine -> LDPC decode with CRC24B stop) -> nr_postDecode (TB CRC)
            openair1/SCHED_NR/phy_procedures_nr_gNB.c:271-300, NR_TRANSPORT/nr_ulsch_demodulation.c:1447, nr_ulsch_decoding.c:320
  transmit: nr_ulsch_encoding-style coding chain (TB CRC, segmentation, LDPC encode, rate match + interleave), scrambling, modulation,
            resource mapping and OFDM modulation -- used here to synthesise a standards-shaped slot for tests and benchmarks.
Channel estimation runs on the DMRS symbol the synthesiser transmits (type 1, port 0, one symbol, no data on it).
torch is used for buffers, index plumbing (resource mapping) and the synthetic channel only."""
import numpy as np
import torch

from . import transport as T
from . import ldpc as _L
from .ldpc import CRC24_A, CRC24_B, PuschChestDesc, PuschRxDesc
from .ofdm import NrOfdmParms


class PuschSlotChain:
    def __init__(self, lib, dl, device, A=235624, N=4096, mu=1, carrier_rb=273, rb_start=0, rb_size=273, nb_rx=4, Qm=6, slot=1, rnti=0x1234, nid=77,
                 ul_freq=3609200000.0, max_iter=8, dmrs_id=55, n_layers=1, transform_precoding=None):
        self.lib, self.dl, self.dev = lib, dl, device
        self.latency_mode = 1            # decoder: a cluster of SMs per code block (one slot alone); the pipelines below run many slots and set 0
        self.P = NrOfdmParms(N, mu, carrier_rb)
        self.N, self.nb_rx, self.Qm, self.slot, self.rnti, self.nid, self.max_iter = N, nb_rx, Qm, slot, rnti, nid, max_iter
        self.rb_start, self.rb_size, self.A, self.nl = rb_start, rb_size, A, n_layers
        self.host_scalars = False
        assert n_layers in (1, 2)
        self.dmrs_pos, self.dmrs_type, self.cdm = 1 << 2, 0, 2                     # one type-1 DMRS symbol, no data on it
        self.seg = T.nr_segmentation(A + 24, 1)
        assert (A + 24 + self.seg["C"] * self.seg["L"]) % (8 * self.seg["C"]) == 0, "pick A like a real TBS: whole bytes per segment"
        self.C, self.K, self.Z, self.F = self.seg["C"], self.seg["K"], self.seg["Z"], self.seg["F"]
        self.seg_crc_type = CRC24_B if self.C > 1 else CRC24_A             # crcType(C, A): a lone segment is checked with the transport block's own CRC
        self.G = T.nr_get_G(rb_size, 14, 12, 1, 0, Qm, n_layers)
        E = [T.nr_get_E(self.G, self.C, Qm, n_layers, r) for r in range(self.C)]
        self.R = T.nr_get_R_ldpc_decoder(0, E[0], 1, self.Z)[0]
        self.E = torch.tensor(E, dtype=torch.int32, device=device)
        self.Eoff = torch.tensor(np.concatenate([[0], np.cumsum(E)[:-1]]), dtype=torch.int32, device=device)
        self.rot = self.P.symbol_rotation(ul_freq)
        self.ts = torch.from_numpy(self.P.timeshift_rotation()).to(device)
        self.desc = PuschRxDesc(N, nb_rx, rb_start, 0, rb_size, self.P.first_carrier_offset, Qm, 0, 14, self.dmrs_pos, self.dmrs_type, self.cdm,
                                0, 14 * N, 14 * N, 1, rnti, nid, n_layers, 0, 0)
        assert lib.pusch_num_llr(self.desc) == self.G
        self.cdescs = [PuschChestDesc(N, nb_rx, slot, 2, p, rb_start, 0, rb_size, self.P.first_carrier_offset, 0, dmrs_id, 14 * N, 14 * N, 1) for p in range(n_layers)]
        self.cdesc = PuschChestDesc(N, nb_rx, slot, 2, 0, rb_start, 0, rb_size, self.P.first_carrier_offset, 0, dmrs_id, 14 * N, 14 * N, n_layers)   # all ports in one call
        # transform precoding (DFT-s-OFDM): (u, v) of the low-PAPR type-1 DMRS; one layer, Qm <= 6, 12 * rb_size one of nr_idft's sizes
        self.tp = transform_precoding
        if self.tp is not None:
            assert n_layers == 1 and Qm <= 6
            seq = lib.lowpapr_sequence(self.tp[0], self.tp[1], 6 * rb_size)
            assert seq is not None, "sequence lengths below 30 are table look-ups the caller provides"
            self._seq_host, self._seq_dev = seq, torch.from_numpy(seq).to(device)
            for cd in self.cdescs:
                cd.set_lowpapr(self._seq_host)                                  # host entry points (pilot generation for the synthesiser)
            self.cdesc.set_lowpapr(self._seq_dev)
            self.desc.transform_precoding = 1
            self._tp_scratch = torch.empty(lib.pusch_tp_scratch_bytes(self.desc), dtype=torch.uint8, device=device)
            self.desc.d_tp_scratch = self._tp_scratch.data_ptr()
        self.est = torch.zeros((n_layers * nb_rx, 14 * N, 2), dtype=torch.int16, device=device)   # ul_ch_estimates[p * nb_rx + aarx]
        self.chest_scratch = torch.empty(lib.pusch_chest_scratch_bytes(self.cdesc), dtype=torch.uint8, device=device)
        self.chest_state = torch.zeros((n_layers, 18), dtype=torch.int32, device=device)
        # receive-side buffers (allocated once, like the reference's per-UE pusch_vars)
        self.rxF = torch.empty((nb_rx, 14 * N, 2), dtype=torch.int16, device=device)
        self.level = torch.zeros(9, dtype=torch.int32, device=device)
        self.llr16 = torch.empty(self.G, dtype=torch.int16, device=device)
        self.harq = torch.zeros((self.C, 66 * self.Z), dtype=torch.int16, device=device)
        self.llr8 = torch.empty((self.C, 68 * self.Z), dtype=torch.int8, device=device)
        self.hard = torch.empty((self.C, 68 * self.Z // 8), dtype=torch.uint8, device=device)
        self.iters = torch.empty(self.C, dtype=torch.int32, device=device)
        self.tb = torch.empty((1, (A + 24) // 8), dtype=torch.uint8, device=device)
        self.tbcrc = torch.empty(1, dtype=torch.int32, device=device)
        self.drx = self.P.desc(slot, nb_rx, self.rot, rx=True)
        # data resource elements in transmission order: symbol-major, sub-carriers from start_re, wrapping at N
        start_re = (self.P.first_carrier_offset + rb_start * 12) % N
        sc = (start_re + np.arange(12 * rb_size)) % N
        syms = [s for s in range(14) if not (self.dmrs_pos >> s) & 1]
        self.re_index = torch.tensor(np.concatenate([s * N + sc for s in syms]), dtype=torch.int64, device=device)
        self._slot_desc = None

    def _c_slot(self, rxdata, use_estimates, est):
        """nrb200_sch_rx_slot_t / nrb200_sch_rx_bufs_t for this chain (include/nrb200_slot.h): the library sequences the slot's launches itself."""
        _L._late_fields()
        d = _L.SchRxSlotDesc()
        C_ = _L.C
        C_.memmove(C_.addressof(d.ofdm), C_.addressof(self.drx), C_.sizeof(self.drx))
        C_.memmove(C_.addressof(d.chest), C_.addressof(self.cdesc), C_.sizeof(self.cdesc))
        C_.memmove(C_.addressof(d.rx), C_.addressof(self.desc), C_.sizeof(self.desc))
        d.rx.d_est_state, d.rx.est_state_ports = 0, 0
        d.rm = self.lib._rmdesc(1, self.Z, self.Qm, 0, self.C, 0, self.F, self.C, 1)
        d.R, d.numMaxIter, d.use_estimates, d.latency_mode = self.R, self.max_iter, 1 if use_estimates else 0, self.latency_mode
        d.crc_len_bits, d.seg_crc_type = self.K - self.F, self.seg_crc_type
        d.A, d.tb_crc_bits, d.seg_payload_bytes = self.A, 24, (self.seg["Kprime"] - self.seg["L"]) // 8
        b = _L.SchRxBufs()
        e = est if use_estimates else self.est
        b.d_rxdata, b.d_timeshift, b.d_rxdataF, b.d_est = rxdata.data_ptr(), self.ts.data_ptr(), self.rxF.data_ptr(), e.data_ptr()
        b.d_chest_scratch, b.d_chest_state, b.d_level, b.d_llr16 = self.chest_scratch.data_ptr(), self.chest_state.data_ptr(), self.level.data_ptr(), self.llr16.data_ptr()
        b.d_E, b.d_Eoff, b.d_harq, b.d_llr8, b.d_hard = self.E.data_ptr(), self.Eoff.data_ptr(), self.harq.data_ptr(), self.llr8.data_ptr(), self.hard.data_ptr()
        b.d_iters, b.d_tb, b.d_tbcrc = self.iters.data_ptr(), self.tb.data_ptr(), self.tbcrc.data_ptr()
        b.harq_stride, b.llr8_stride, b.hard_stride = self.harq.shape[1], self.llr8.shape[1], self.hard.shape[1]
        return d, b

    # ------------------------------------------------------------------ synthesis (not timed)
    def synthesize(self, seed=1, snr_db=30.0, h_amp=724.0, tx_amp=724):
        """Returns (payload bytes uint8[A/8], rxdata int16 [nb_rx, samples_per_frame, 2], ul_ch_estimates int16 [nb_rx, 14*N, 2])."""
        lib, dl, dev, N = self.lib, self.dl, self.dev, self.N
        rng = np.random.default_rng(seed)
        payload = rng.integers(0, 256, size=self.A // 8, dtype=np.uint8)
        crc = int(lib.crc_batch_host(0, payload[None, :], self.A)[0]) >> 8                  # crc24a
        tb = np.concatenate([payload, np.array([(crc >> 16) & 255, (crc >> 8) & 255, crc & 255], np.uint8)])
        segs, _ = T.segment_transport_block(lib, tb, 1)
        cw = lib.encode_batch_torch(1, self.Z, self.K, torch.from_numpy(segs).to(dev))
        f = torch.empty(self.G, dtype=torch.uint8, device=dev)
        lib.rm_tx_torch(1, self.Z, self.Qm, 0, self.C, 0, self.F, cw, self.E, self.Eoff, f)
        words = torch.zeros((self.G + 31) // 32 + 1, dtype=torch.int32, device=dev)
        lib.scramble_torch(f, 0, self.nid, self.rnti, words)
        sym = torch.empty((self.G // self.Qm, 2), dtype=torch.int16, device=dev)
        lib.modulate_torch(words, self.G, self.Qm, sym)
        x = (sym.to(torch.int32) * tx_amp) >> 15                                           # the amp scaling of the TX resource mapper
        nl = self.nl
        xl = x.reshape(-1, nl, 2)                                                          # layer mapping: x^(l)(i) = d(nl * i + l)
        g = torch.Generator(device=dev); g.manual_seed(seed)
        # flat nb_rx x nl channel: random phases, layer 1 orthogonal-ish to layer 0 across antenna pairs
        ph = torch.rand((self.nb_rx, nl), generator=g, device=dev) * 6.2831853
        if nl == 2:
            ph[:, 1] = ph[:, 0] + 3.14159265 * (torch.arange(self.nb_rx, device=dev) % 2) + 0.3 * torch.arange(self.nb_rx, device=dev)
        hi = torch.round(torch.stack([torch.cos(ph), torch.sin(ph)], dim=2) * h_amp).to(torch.int32)      # [rx][layer][re/im]
        # Genie estimate in the reference's convention: rxdataF = h_est * x_unit with x_unit the unit-energy constellation (the estimator
        # divides by unit-amplitude pilots, so the transmit amplitude is part of h_est).  Here rxdataF = hi * x / 1024 and
        # x = x_unit * 23170 * tx_amp / 32768.
        unit = 23170.0 * tx_amp / 32768.0
        he = torch.round(hi.to(torch.float32) * (unit / 1024.0)).to(torch.int16)
        est = torch.zeros((nl * self.nb_rx, 14 * N, 2), dtype=torch.int16, device=dev)
        dm = 2
        for l in range(nl):
            est[l * self.nb_rx:(l + 1) * self.nb_rx, dm * N:dm * N + 12 * self.rb_size, :] = he[:, l, None, :]
        sigma = unit * (h_amp / 1024.0) * 10.0 ** (-snr_db / 20.0) * 0.70711
        grid = torch.zeros((self.nb_rx, 14 * N, 2), dtype=torch.float32, device=dev)
        # DMRS (type 1, ports 0..nl-1 on every second sub-carrier of symbol 2, CDM-separated by the w_f cover) = conj of the receiver's pilot
        # tables at the data's unit amplitude
        pils = [torch.from_numpy(lib.pusch_dmrs_pilots(cd).reshape(-1, 2).astype(np.float32)).to(dev) * (unit / 32767.0) for cd in self.cdescs]
        k0 = (self.P.first_carrier_offset + self.rb_start * 12) % N
        dm_index = torch.tensor(2 * N + (k0 + 2 * np.arange(6 * self.rb_size)) % N, dtype=torch.int64, device=dev)
        hf = hi.to(torch.float32) / 1024.0
        xf = xl.to(torch.float32)
        if self.tp is not None:
            # what the UE's transform precoder does to every symbol's M modulation symbols (38.211 6.3.1.4): X = DFT_M(x) / sqrt(M)
            M = 12 * self.rb_size
            xc = torch.view_as_complex(xf[:, 0, :].reshape(-1, M, 2).contiguous())
            xf = torch.view_as_real(torch.fft.fft(xc, dim=1) / (M ** 0.5)).reshape(-1, 1, 2).contiguous()
        for a in range(self.nb_rx):
            yr = sum(hf[a, l, 0] * xf[:, l, 0] - hf[a, l, 1] * xf[:, l, 1] for l in range(nl))
            yi = sum(hf[a, l, 0] * xf[:, l, 1] + hf[a, l, 1] * xf[:, l, 0] for l in range(nl))
            y = torch.stack([yr, yi], dim=1) + sigma * torch.randn((xf.shape[0], 2), generator=g, device=dev)
            grid[a].index_copy_(0, self.re_index, y)
            dr = sum(hf[a, l, 0] * pils[l][:, 0] + hf[a, l, 1] * pils[l][:, 1] for l in range(nl))       # sum_l h_l * conj(pil_l)
            di = sum(hf[a, l, 1] * pils[l][:, 0] - hf[a, l, 0] * pils[l][:, 1] for l in range(nl))
            grid[a].index_copy_(0, dm_index, torch.stack([dr, di], dim=1) + sigma * torch.randn((pils[0].shape[0], 2), generator=g, device=dev))
        gridF = torch.clamp(torch.round(grid), -32768, 32767).to(torch.int16).contiguous()
        # to the time domain with the library's own modulator (phase pre-compensation on, as a UE would transmit)
        dtx = self.P.desc(self.slot, self.nb_rx, self.rot)
        t = torch.zeros((self.nb_rx, dtx.t_stride, 2), dtype=torch.int16, device=dev)
        dl.ofdm_mod_slot_torch(dtx, gridF, t)
        rxdata = torch.zeros((self.nb_rx, self.P.samples_per_frame, 2), dtype=torch.int16, device=dev)
        ss = self.P.slot_timestamp(self.slot)
        rxdata[:, ss:ss + dtx.t_stride] = t
        torch.cuda.synchronize()
        return payload, rxdata, est

    # ------------------------------------------------------------------ receive chain (the timed part)
    def receive(self, rxdata, est=None, staged=False):
        """est=None: estimate the channel from the DMRS symbol (the normal path); otherwise use the caller's ul_ch_estimates.
        The slot is ONE library call (nrb200_sch_slot_rx_dev); staged=True issues the same stages one entry point at a time from here (the round-1 form, kept
        for the test that compares the two)."""
        lib, dl = self.lib, self.dl
        if not staged and not self.host_scalars:
            d, b = self._c_slot(rxdata, est is not None, est)
            self._keep = (d, b)
            lib.sch_slot_rx_torch(d, b, self.dev)
            return self.tb, self.iters, self.tbcrc
        dl.ofdm_demod_slot_torch(self.drx, rxdata, self.ts, self.rxF)
        if est is None:
            lib.pusch_chest_torch(self.cdesc, self.rxF, self.est, self.chest_scratch, self.chest_state)   # every DMRS port of the PDU (:1473-1486)
            est = self.est
            if self.nl == 2:
                # the MMSE receiver needs two scalars of the estimator, max_ch and nvar (:1470-1512): read on the device from the estimator's state
                # (host_scalars = True keeps the older path through the descriptor, one device-to-host read per slot)
                if self.host_scalars:
                    st = self.chest_state[:, :2].cpu().numpy()
                    self.desc.d_est_state = 0
                    self.desc.max_ch = int(st[:, 0].max())
                    self.desc.noise_var = int(int(st[:, 1].astype(np.int64).sum()) // (14 * self.nl * self.nb_rx))
                else:
                    self.desc.d_est_state, self.desc.est_state_ports = self.chest_state.data_ptr(), self.nl
        lib.pusch_inner_rx_torch(self.desc, self.rxF, est, self.llr16, level=self.level)
        lib.rm_rx_torch(1, self.Z, self.Qm, 0, self.C, 0, self.F, self.llr16, self.E, self.Eoff, self.harq, self.llr8, clear=1)
        lib.decode_batch_torch(1, self.Z, self.R, self.max_iter, self.llr8, use_crc=1, crc_len_bits=self.K - self.F, crc_type=self.seg_crc_type,
                               out=self.hard, iters=self.iters, latency_mode=self.latency_mode)
        nbytes = (self.seg["Kprime"] - self.seg["L"]) // 8
        self.tb.view(-1).copy_(self.hard[:, :nbytes].reshape(-1))                             # nr_postDecode: concatenate the segments
        lib.crc_batch_torch(0, self.tb, self.A + 24, out=self.tbcrc)                       # CRC over payload + CRC24A == 0 when intact
        return self.tb, self.iters, self.tbcrc


class PuschSlotPipeline:
    """K PUSCH slots in flight on one GPU, the uplink counterpart of dl_slot_chain.PdschSlotPipeline: K independent PuschSlotChain instances (own buffers,
    own stream, own received samples), each slot's receive launches captured once into a CUDA graph and replayed.  One slot alone is latency bound (11 dependent
    launches, 28 code blocks on 148 SMs); a gNB serves several users and carriers, whose slots are independent.  Two layers work the same way: the MMSE
    receiver reads the estimator's max_ch / nvar from device memory (nrb200_pusch_rx_t.d_est_state), so the whole slot is stream ordered and capturable."""

    def __init__(self, lib, dl, device, n_inflight, use_graphs=True, seed0=200, **cfg):
        self.K, self.dev = n_inflight, device
        self.chains, self.graphs, self.streams, self.rx, self.payload, self.h_rx, self.h_tb = [], [], [], [], [], [], []
        for k in range(n_inflight):
            st = torch.cuda.Stream(device=device)
            with torch.cuda.stream(st):
                ch = PuschSlotChain(lib, dl, device, **cfg)
                ch.latency_mode = 0                                  # many slots in flight: one CTA per code block spends the fewest SM-cycles
                payload, rxdata, _ = ch.synthesize(seed=seed0 + k, snr_db=30.0)
                ch.receive(rxdata)                                   # warm-up outside the capture
                st.synchronize()
                g = None
                if use_graphs:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=st):
                        ch.receive(rxdata)
            self.chains.append(ch); self.graphs.append(g); self.streams.append(st); self.rx.append(rxdata); self.payload.append(payload)
            ss, ns = ch.P.slot_timestamp(ch.slot), ch.P.samples_per_slot(ch.slot)
            self.slot_span = (ss, ss + ns)
            self.h_rx.append(rxdata[:, ss:ss + ns].contiguous().cpu().pin_memory())       # the slot's samples only (983 KB at 4 rx)
            self.h_tb.append(torch.empty_like(ch.tb, device="cpu").pin_memory())
        torch.cuda.synchronize(device)

    def round(self, e2e=False):
        for k in range(self.K):
            with torch.cuda.stream(self.streams[k]):
                if e2e:
                    self.rx[k][:, self.slot_span[0]:self.slot_span[1]].copy_(self.h_rx[k], non_blocking=True)   # the slot's time-domain samples of every rx antenna
                if self.graphs[k] is not None:
                    self.graphs[k].replay()
                else:
                    self.chains[k].receive(self.rx[k])
                if e2e:
                    self.h_tb[k].copy_(self.chains[k].tb, non_blocking=True)

    def timed_rounds(self, n_rounds, e2e=False, warm=3):
        """CUDA-event time (ms) of n_rounds x K slots; events on the current stream, every slot stream forks from it and joins back."""
        for _ in range(warm):
            self.round(e2e)
        torch.cuda.synchronize(self.dev)
        cur = torch.cuda.current_stream(self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for st in self.streams:
            st.wait_event(e0)
        for _ in range(n_rounds):
            self.round(e2e)
        for st in self.streams:
            ev = torch.cuda.Event(); ev.record(st); cur.wait_event(ev)
        e1.record(cur)
        torch.cuda.synchronize(self.dev)
        return e0.elapsed_time(e1)

    def check(self, host=False):
        ok = []
        for k, ch in enumerate(self.chains):
            tb = self.h_tb[k] if host else ch.tb.cpu()
            want = torch.from_numpy(self.payload[k])
            ok.append(bool((ch.iters <= ch.max_iter).all()) and int(ch.tbcrc.cpu()[0]) == 0 and bool((tb.view(-1)[:want.numel()] == want).all()))
        return ok
