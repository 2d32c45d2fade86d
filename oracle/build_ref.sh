#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* OpenAirInterface reference sources,
# in place from /root/reference, into shared objects under oracle/_ref/ (git-ignored).
# Nothing from the reference tree is copied into this repository; the only generated text
# is a simde->native intrinsic alias header (simde is absent in this image; on x86 simde is a
# 1:1 wrapper of the native intrinsics, reference CMakeLists.txt:124-133) and the reference's
# own code-generator output (nrLDPC_tools/generator_*), both written to oracle/_ref/.
#
# Products (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm
# may load them):
#   oracle/_ref/libref_ldpc_dec.so      LDPCinit/LDPCshutdown/LDPCdecoder  (nrLDPC_decoder.c, AVX2)
#   oracle/_ref/libref_ldpc_dec512.so   same, -march=native (AVX512 generated kernels) [optional]
#   oracle/_ref/libref_ldpc_enc.so      LDPCencoder (ldpc_encoder_optim8segmulti.c, default libldpc.so)
#   oracle/_ref/libref_ldpc_enc_orig.so LDPCencoder (ldpc_encoder.c, scalar "_orig")
#   oracle/_ref/libref_dfts.so          dft/idft/dfts_autoinit (oai_dfts.c)
#   oracle/_ref/libref_coding.so        crc_byte.c + nr_rate_matching.c + nr_segmentation.c
#   oracle/_ref/libref_mod.so           nr_modulation.c + nr_gen_mod_table.c (QAM mapper)
#   oracle/_ref/libref_llr.so           nr_ulsch_llr_computation.c (PUSCH max-log LLRs)
set -euo pipefail
R=${OAI_REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
W=$HERE/_ref
if [ ! -d "$R/openair1" ]; then echo "reference tree not present at $R: keeping prebuilt oracle/_ref" >&2; exit 0; fi
mkdir -p $W/shim/simde/x86 $W/shim/simde/arm $W/gen/{cnProc,bnProc,bnProcPc,cnProc_avx512,bnProc_avx512,bnProcPc_avx512}
# 1) simde -> native alias header
grep -rhoE '\bsimde_[_a-zA-Z0-9]+|\bsimde__m[0-9a-z]+|\bSIMDE_[A-Z_0-9]+' $R/openair1 $R/common | sort -u > $W/tokens.txt
{ echo '#pragma once'; echo '#include <immintrin.h>'; echo '#include <mmintrin.h>';
  while read t; do case "$t" in
    # AVX512VL-only spellings that simde emulates on AVX2: same unaligned 128/256-bit move
    simde_mm256_loadu_epi32|simde_mm256_loadu_epi16|simde_mm256_loadu_epi8) echo "#define $t(p) _mm256_loadu_si256((const __m256i *)(p))";;
    simde_mm_loadu_epi32|simde_mm_loadu_epi16|simde_mm_loadu_epi8) echo "#define $t(p) _mm_loadu_si128((const __m128i *)(p))";;
    simde_mm256_storeu_epi32|simde_mm256_storeu_epi16|simde_mm256_storeu_epi8) echo "#define $t(p, v) _mm256_storeu_si256((__m256i *)(p), (v))";;
    simde_mm_storeu_epi32|simde_mm_storeu_epi16|simde_mm_storeu_epi8) echo "#define $t(p, v) _mm_storeu_si128((__m128i *)(p), (v))";;
    simde_*) echo "#define $t ${t#simde}";; SIMDE_MM_SHUFFLE) echo "#define $t _MM_SHUFFLE";; esac; done < $W/tokens.txt; } > $W/shim/simde/x86/shim_all.h
for f in mmx sse sse2 sse3 ssse3 sse4.1 sse4.2 avx2 fma clmul avx512; do echo '#include "shim_all.h"' > $W/shim/simde/x86/$f.h; done
echo '#include "x86/shim_all.h"' > $W/shim/simde/simde-common.h; echo '#pragma once' > $W/shim/simde/arm/neon.h
INC="-I$W/shim -I$W/gen -I$R/openair1/PHY/CODING/nrLDPC_decoder -I$R/openair1 -I$R -I$R/common/utils -I$R/common/utils/LOG -I$R/common/utils/T \
 -I$R/openair2/COMMON -I$R/nfapi/open-nFAPI/nfapi/public_inc -I$R/openair2 -I$R/openair1/PHY -I$R/common -I$R/radio/COMMON -I$R/executables \
 -I$R/openair2/NR_UE_PHY_INTERFACE -I$R/openair2/NR_PHY_INTERFACE -I$R/openair2/PHY_INTERFACE -I$R/openair3/COMMON -I$R/openair3"
DEFS="-DMAX_NUM_CCs=1 -DNB_ANTENNAS_RX=4 -DNB_ANTENNAS_TX=4 -DNUMBER_OF_UE_MAX_NB_IoT=16"
# 2) the reference's own generated decoder headers
T=$R/openair1/PHY/CODING/nrLDPC_decoder/nrLDPC_tools
if [ ! -f $W/gen/.done ]; then
  ( cd $W/gen
    gcc -O1 -mavx2 $INC -DCODEGEN $T/generator_cnProc/{cnProc_gen_BG1_avx2.c,cnProc_gen_BG2_avx2.c,main.c} -o cn_gen && ./cn_gen .
    gcc -O1 -mavx2 $INC -DCODEGEN $T/generator_bnProc/{bnProc_gen_BG1_avx2.c,bnProc_gen_BG2_avx2.c,bnProcPc_gen_BG1_avx2.c,bnProcPc_gen_BG2_avx2.c,main.c} -o bn_gen && ./bn_gen .
    gcc -O1 -mavx512bw $INC -DCODEGEN $T/generator_cnProc_avx512/*.c -o cn512 && ./cn512 . || true
    gcc -O1 -mavx512bw $INC -DCODEGEN $T/generator_bnProc_avx512/*.c -o bn512 && ./bn512 . || true
    touch .done )
fi
# 3) libraries
cd $W
F="-O3 -mavx2 -mno-avx512f -fPIC -shared -w"
gcc $F $INC $DEFS $HERE/ref_stubs.c $R/openair1/PHY/CODING/nrLDPC_decoder/nrLDPC_decoder.c              -o libref_ldpc_dec.so
gcc $F $INC $DEFS $HERE/ref_stubs.c $R/openair1/PHY/CODING/nrLDPC_encoder/ldpc_encoder_optim8segmulti.c -o libref_ldpc_enc.so
gcc $F $INC $DEFS $HERE/ref_stubs.c $R/openair1/PHY/CODING/nrLDPC_encoder/ldpc_encoder.c                -o libref_ldpc_enc_orig.so
gcc $F $INC $DEFS $HERE/ref_stubs.c $R/openair1/PHY/TOOLS/oai_dfts.c -lm                                -o libref_dfts.so
gcc -O3 -march=native -fPIC -shared -w $INC $DEFS $HERE/ref_stubs.c $R/openair1/PHY/CODING/nrLDPC_decoder/nrLDPC_decoder.c -o libref_ldpc_dec512.so || echo "avx512 variant skipped"
gcc $F -mpclmul $INC $DEFS $HERE/ref_stubs.c $R/openair1/PHY/CODING/crc_byte.c $R/openair1/PHY/CODING/nr_rate_matching.c \
    $R/openair1/PHY/CODING/nr_segmentation.c -o libref_coding.so || echo "libref_coding.so: FAILED (see DESIGN.md)"
gcc $F $INC $DEFS $HERE/ref_stubs.c $R/openair1/PHY/NR_TRANSPORT/nr_ulsch_llr_computation.c $R/openair1/PHY/TOOLS/simde_operations.c -o libref_llr.so || echo "libref_llr.so: FAILED"
gcc $F $INC $DEFS $HERE/ref_stubs.c $HERE/ref_stubs_mod.c $R/openair1/PHY/MODULATION/nr_modulation.c $R/openair1/PHY/NR_REFSIG/nr_gen_mod_table.c $R/openair1/PHY/NR_TRANSPORT/nr_scrambling.c $R/openair1/PHY/NR_REFSIG/scrambling_luts.c -lm -o libref_mod.so || echo "libref_mod.so: FAILED"
# slot-level OFDM front end; dft/idft stay function pointers bound at run time to libref_dfts.so (ref_harness_ofdm.c)
gcc $F $INC $DEFS $HERE/ref_stubs.c $HERE/ref_harness_ofdm.c $R/openair1/PHY/MODULATION/ofdm_mod.c $R/openair1/PHY/MODULATION/slot_fep_nr.c \
    $R/openair1/PHY/TOOLS/cmult_sv.c $R/openair1/PHY/TOOLS/cmult_vv.c $R/openair1/PHY/MODULATION/nr_modulation.c $R/openair1/PHY/NR_REFSIG/nr_gen_mod_table.c \
    -lm -ldl -o libref_ofdm.so || echo "libref_ofdm.so: FAILED"
# PUSCH inner receiver: ref_harness_pusch.c textually includes nr_ulsch_demodulation.c (its per-symbol functions are static)
gcc $F $INC $DEFS $HERE/ref_stubs.c $HERE/ref_stubs_pusch.c $HERE/ref_harness_pusch.c $R/openair1/PHY/NR_TRANSPORT/nr_ulsch_llr_computation.c \
    $R/openair1/PHY/TOOLS/simde_operations.c $R/openair1/PHY/TOOLS/log2_approx.c $R/openair1/PHY/NR_ESTIMATION/nr_freq_equalization.c -lm -ldl -o libref_pusch.so || echo "libref_pusch.so: FAILED"
# PUSCH channel estimation (nr_common.c needs <limits.h> for UINT_MAX; dft/idft are bound to libref_dfts.so at run time)
gcc $F -include limits.h $INC $DEFS $HERE/ref_stubs.c $HERE/ref_stubs_chest.c $HERE/ref_harness_chest.c $R/openair1/PHY/NR_ESTIMATION/nr_ul_channel_estimation.c \
    $R/openair1/PHY/NR_REFSIG/nr_dmrs_rx.c $R/openair1/PHY/NR_REFSIG/nr_gold.c $R/common/utils/nr/nr_common.c $R/openair1/PHY/TOOLS/cmult_sv.c \
    $R/openair1/PHY/TOOLS/log2_approx.c $R/openair1/PHY/NR_REFSIG/ul_ref_seq_nr.c -lm -ldl -o libref_chest.so || echo "libref_chest.so: FAILED"
# UE-side PDSCH channel estimation
gcc $F -include limits.h $INC $DEFS $HERE/ref_stubs.c $HERE/ref_stubs_uechest.c $HERE/ref_harness_uechest.c $R/openair1/PHY/NR_UE_ESTIMATION/nr_dl_channel_estimation.c \
    $R/openair1/PHY/NR_REFSIG/nr_dmrs_rx.c $R/openair1/PHY/NR_REFSIG/nr_gold_ue.c $R/openair1/PHY/NR_REFSIG/dmrs_nr.c $R/openair1/PHY/NR_TRANSPORT/nr_sch_dmrs.c $R/openair1/PHY/NR_REFSIG/nr_gen_mod_table.c \
    $R/common/utils/nr/nr_common.c $R/openair1/PHY/TOOLS/cmult_sv.c $R/openair1/PHY/TOOLS/cmult_vv.c $R/openair1/PHY/MODULATION/slot_fep_nr.c $R/openair1/PHY/TOOLS/log2_approx.c -lm -ldl -Wl,--no-undefined -o libref_uechest.so || echo "libref_uechest.so: FAILED"
# UE-side PDSCH receiver: the real nr_rx_pdsch, symbol by symbol
gcc $F -include limits.h $INC $DEFS $HERE/ref_stubs.c $HERE/ref_stubs_pdsch.c $HERE/ref_harness_pdsch.c $R/openair1/PHY/NR_UE_TRANSPORT/nr_dlsch_demodulation.c \
    $R/openair1/PHY/NR_UE_TRANSPORT/nr_dlsch_llr_computation.c $R/openair1/PHY/NR_REFSIG/dmrs_nr.c $R/openair1/PHY/TOOLS/log2_approx.c -lm -o libref_pdsch.so || echo "libref_pdsch.so: FAILED"
# the same receiver harness with PT-RS: the real nr_pdsch_ptrs_processing (nr_dl_channel_estimation.c) + ptrs_nr.c linked in instead of the abort stub
gcc $F -DREFH_PTRS -include limits.h $INC $DEFS $HERE/ref_stubs.c $HERE/ref_stubs_pdsch_ptrs.c $HERE/ref_harness_pdsch.c $R/openair1/PHY/NR_UE_TRANSPORT/nr_dlsch_demodulation.c \
    $R/openair1/PHY/NR_UE_TRANSPORT/nr_dlsch_llr_computation.c $R/openair1/PHY/NR_UE_ESTIMATION/nr_dl_channel_estimation.c $R/openair1/PHY/NR_REFSIG/ptrs_nr.c \
    $R/openair1/PHY/NR_REFSIG/nr_dmrs_rx.c $R/openair1/PHY/NR_REFSIG/nr_gold_ue.c $R/openair1/PHY/NR_REFSIG/dmrs_nr.c $R/openair1/PHY/NR_TRANSPORT/nr_sch_dmrs.c $R/openair1/PHY/NR_REFSIG/nr_gen_mod_table.c \
    $R/common/utils/nr/nr_common.c $R/openair1/PHY/TOOLS/cmult_sv.c $R/openair1/PHY/TOOLS/cmult_vv.c $R/openair1/PHY/MODULATION/slot_fep_nr.c $R/openair1/PHY/TOOLS/log2_approx.c \
    -lm -ldl -Wl,--no-undefined -o libref_pdsch_ptrs.so || echo "libref_pdsch_ptrs.so: FAILED"
# rfsimulator channel application: the real rxAddInput, noise draws and the debug line's signal_energy supplied by the harness
gcc $F -include limits.h $INC $DEFS $HERE/ref_stubs.c $HERE/ref_harness_rfsim.c $R/radio/rfsimulator/apply_channelmod.c -lm -Wl,--no-undefined -o libref_rfsim.so || echo "libref_rfsim.so: FAILED"
# gNB PRACH detector: the real rx_nr_prach + compute_nr_prach_seq + dB_fixed_times10 (idft bound to libref_dfts.so at run time)
gcc $F -include limits.h $INC $DEFS $HERE/ref_stubs.c $HERE/ref_harness_prach.c $R/openair1/PHY/NR_TRANSPORT/nr_prach.c $R/openair1/PHY/NR_TRANSPORT/nr_prach_common.c \
    $R/openair1/PHY/TOOLS/dB_routines.c $R/openair1/PHY/TOOLS/signal_energy.c -lm -ldl -Wl,--no-undefined -o libref_prach.so || echo "libref_prach.so: FAILED"
# gNB-side PDSCH transmitter after the encoder: the real nr_generate_pdsch with nr_dlsch_encoding replaced by the harness (bits in)
gcc $F -include limits.h $INC $DEFS $HERE/ref_stubs.c $HERE/ref_stubs_pdschtx.c $HERE/ref_harness_pdschtx.c $R/openair1/PHY/NR_TRANSPORT/nr_dlsch.c \
    $R/openair1/PHY/NR_REFSIG/nr_gold.c $R/openair1/PHY/NR_TRANSPORT/nr_sch_dmrs.c $R/openair1/PHY/NR_REFSIG/dmrs_nr.c $R/openair1/PHY/NR_REFSIG/ptrs_nr.c \
    $R/common/utils/nr/nr_common.c $R/openair1/PHY/MODULATION/nr_modulation.c $R/openair1/PHY/NR_REFSIG/nr_gen_mod_table.c $R/openair1/PHY/NR_TRANSPORT/nr_scrambling.c \
    $R/openair1/PHY/NR_REFSIG/scrambling_luts.c -lm -o libref_pdschtx.so || echo "libref_pdschtx.so: FAILED"
ls -la $W/*.so
# the reference's side of the LDPC loader boundary (OAI's own types; dlopens the library under test at run time)
gcc $F $INC $DEFS $HERE/ref_stubs.c $HERE/ref_harness_loader.c -ldl -lpthread -o libref_loader.so || echo "libref_loader.so: FAILED"
# OAI's dft_size_idx_t / idft_size_idx_t enumerator order (pins the size index the dft()/idft() drop-in receives)
gcc $F $INC $DEFS $HERE/ref_stubs.c $HERE/ref_harness_dftidx.c -o libref_dftidx.so || echo "libref_dftidx.so: FAILED"
# ---- the reference's OWN physim executable: ldpctest (openair1/PHY/CODING/TESTBENCH/ldpctest.c) with the reference's module loader
#      (load_module_shlib.c), command-line-only config module, logging and noise generators, linked like CMakeLists.txt:2205-2217 does, plus the
#      two LDPC modules it always loads (CMakeLists.txt:816-844): libldpc_orig.so (ldpc_encoder.c) and libldpc.so (ldpc_encoder_optim8segmulti.c),
#      both with nrLDPC_decoder.c.  `ldpctest -v _b200` then dlopens libldpc_b200.so through the unmodified loader (tests/test_gpu_ldpctest.py).
mkdir -p $W/oai_libs $W/ldpctest_obj
LT_DEFS="$DEFS -DPACKAGE_VERSION=\"oracle-build\" -DT_TRACER=0"
LT_SRCS="openair1/PHY/CODING/TESTBENCH/ldpctest.c openair1/PHY/CODING/nrLDPC_load.c common/utils/load_module_shlib.c common/config/config_load_configmodule.c
 common/config/config_userapi.c common/config/config_cmdline.c common/config/config_common.c common/utils/LOG/log.c common/utils/time_meas.c
 openair1/SIMULATION/TOOLS/rangen_double.c openair1/SIMULATION/TOOLS/taus.c common/utils/utils.c"
ok=1
for f in $LT_SRCS; do gcc -O2 -mavx2 -w -c $INC $LT_DEFS $R/$f -o $W/ldpctest_obj/$(basename $f .c).o || { echo "ldpctest: FAILED compiling $f"; ok=0; }; done
if [ $ok = 1 ]; then
  gcc -o $W/ldpctest $W/ldpctest_obj/*.o -lm -lpthread -ldl -rdynamic || echo "ldpctest: FAILED linking"
  gcc $F $INC $DEFS $R/openair1/PHY/CODING/nrLDPC_decoder/nrLDPC_decoder.c $R/openair1/PHY/CODING/nrLDPC_encoder/ldpc_encoder.c -o $W/oai_libs/libldpc_orig.so
  gcc $F $INC $DEFS $R/openair1/PHY/CODING/nrLDPC_decoder/nrLDPC_decoder.c $R/openair1/PHY/CODING/nrLDPC_encoder/ldpc_encoder_optim8segmulti.c -o $W/oai_libs/libldpc.so
  ls -la $W/ldpctest $W/oai_libs
fi
# ---- the reference's nr_ulsch_decoding (segment jobs on the thread pool, rate recovery, one LDPCdecoder call per segment) behind a caller harness that
#      collects the results like phy_procedures_gNB_uespec_RX does; ldpc_interface is bound at run time to oai_libs/libldpc.so (the reference CPU decoder)
gcc $F -mpclmul -D_GNU_SOURCE -include limits.h $INC $DEFS $HERE/ref_stubs.c $HERE/ref_stubs_ulsch.c $HERE/ref_harness_ulsch.c $R/openair1/PHY/NR_TRANSPORT/nr_ulsch_decoding.c \
    $R/openair1/PHY/CODING/nr_segmentation.c $R/openair1/PHY/CODING/nr_rate_matching.c $R/openair1/PHY/CODING/crc_byte.c $R/openair1/PHY/NR_TRANSPORT/nr_tbs_tools.c \
    $R/openair1/PHY/TOOLS/dB_routines.c $R/common/utils/threadPool/thread-pool.c -lm -ldl -lpthread -Wl,--no-undefined -o libref_ulsch.so || echo "libref_ulsch.so: FAILED"
