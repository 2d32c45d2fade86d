#!/usr/bin/env bash
# Builds the OAI-side interposers of integration/ against the reference's own headers (they define OAI symbols, so they need OAI's types):
#   integration/_build/libnrb200_shim_chest.so      LD_PRELOAD-able nr_pusch_channel_estimation -> libldpc_b200.so
#   oracle/_ref/libshimtest_chest.so                the reference-side caller harness (oracle/ref_harness_chest.c, test infrastructure) linked against the
#                                                   interposer INSTEAD of nr_ul_channel_estimation.c: what tests/test_gpu_interpose.py drives on the GPU box
# Needs oracle/build_ref.sh to have run (simde alias headers) and openairinterface5g_b200/libldpc_b200.so to exist.  Where /root/reference is absent the
# prebuilt files are kept (they travel to the GPU box like every other built artefact).
set -euo pipefail
R=${OAI_REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(dirname "$HERE")
W=$ROOT/oracle/_ref
if [ ! -d "$R/openair1" ]; then echo "reference tree not present at $R: keeping prebuilt interposers" >&2; exit 0; fi
mkdir -p $HERE/_build
INC="-I$W/shim -I$R/openair1 -I$R -I$R/common/utils -I$R/common/utils/LOG -I$R/common/utils/T \
 -I$R/openair2/COMMON -I$R/nfapi/open-nFAPI/nfapi/public_inc -I$R/openair2 -I$R/openair1/PHY -I$R/common -I$R/radio/COMMON -I$R/executables \
 -I$R/openair2/NR_UE_PHY_INTERFACE -I$R/openair2/NR_PHY_INTERFACE -I$R/openair2/PHY_INTERFACE -I$R/openair3/COMMON -I$R/openair3 -I$ROOT/include"
DEFS="-DMAX_NUM_CCs=1 -DNB_ANTENNAS_RX=4 -DNB_ANTENNAS_TX=4 -DNUMBER_OF_UE_MAX_NB_IoT=16"
F="-O2 -mavx2 -mno-avx512f -fPIC -shared -w -include limits.h"
LIB="-L$ROOT/openairinterface5g_b200 -l:libldpc_b200.so"
gcc $F $INC $DEFS $HERE/oai_shim_pusch_chest.c $LIB -Wl,-rpath,'$ORIGIN/../../openairinterface5g_b200' -o $HERE/_build/libnrb200_shim_chest.so
gcc $F $INC $DEFS $ROOT/oracle/ref_stubs.c $ROOT/oracle/ref_stubs_chest.c $ROOT/oracle/ref_harness_chest.c $HERE/oai_shim_pusch_chest.c \
    $R/openair1/PHY/NR_REFSIG/nr_dmrs_rx.c $R/openair1/PHY/NR_REFSIG/nr_gold.c $R/common/utils/nr/nr_common.c $R/openair1/PHY/TOOLS/cmult_sv.c \
    $R/openair1/PHY/TOOLS/log2_approx.c $R/openair1/PHY/NR_REFSIG/ul_ref_seq_nr.c $LIB -Wl,-rpath,'$ORIGIN/../../openairinterface5g_b200' -lm -ldl -o $W/libshimtest_chest.so
# the UE-side twin: nr_pdsch_channel_estimation
gcc $F $INC $DEFS $HERE/oai_shim_pdsch_chest.c $LIB -Wl,-rpath,'$ORIGIN/../../openairinterface5g_b200' -o $HERE/_build/libnrb200_shim_uechest.so
gcc $F $INC $DEFS $ROOT/oracle/ref_stubs.c $ROOT/oracle/ref_stubs_uechest.c $ROOT/oracle/ref_harness_uechest.c $HERE/oai_shim_pdsch_chest.c \
    $R/openair1/PHY/NR_REFSIG/nr_dmrs_rx.c $R/openair1/PHY/NR_REFSIG/nr_gold_ue.c $R/openair1/PHY/NR_REFSIG/dmrs_nr.c $R/openair1/PHY/NR_TRANSPORT/nr_sch_dmrs.c \
    $R/openair1/PHY/NR_REFSIG/nr_gen_mod_table.c $R/common/utils/nr/nr_common.c $R/openair1/PHY/TOOLS/cmult_sv.c $R/openair1/PHY/TOOLS/cmult_vv.c \
    $R/openair1/PHY/MODULATION/slot_fep_nr.c $R/openair1/PHY/TOOLS/log2_approx.c $LIB -Wl,-rpath,'$ORIGIN/../../openairinterface5g_b200' -lm -ldl -o $W/libshimtest_uechest.so
ls -la $HERE/_build/*.so $W/libshimtest_chest.so $W/libshimtest_uechest.so
# the UE's PDSCH receiver: nr_rx_pdsch (the caller harness drives it symbol by symbol like nr_ue_pdsch_procedures)
gcc $F $INC $DEFS $HERE/oai_shim_rx_pdsch.c $LIB -Wl,-rpath,'$ORIGIN/../../openairinterface5g_b200' -o $HERE/_build/libnrb200_shim_rx_pdsch.so
gcc $F -DREFH_PTRS $INC $DEFS $ROOT/oracle/ref_stubs.c $ROOT/oracle/ref_stubs_pdsch.c $ROOT/oracle/ref_harness_pdsch.c $HERE/oai_shim_rx_pdsch.c \
    $R/openair1/PHY/NR_REFSIG/dmrs_nr.c $R/openair1/PHY/NR_REFSIG/nr_gold_ue.c $R/openair1/PHY/TOOLS/log2_approx.c $LIB -Wl,-rpath,'$ORIGIN/../../openairinterface5g_b200' -lm -o $W/libshimtest_pdsch.so
ls -la $HERE/_build/libnrb200_shim_rx_pdsch.so $W/libshimtest_pdsch.so
# the gNB's PUSCH receiver: nr_rx_pusch_tp (+ the estimator it calls by name), driven by a caller harness that fills PHY_VARS_gNB like phy_init_nr_gNB does
gcc $F $INC $DEFS $HERE/oai_shim_rx_pusch.c $LIB -Wl,-rpath,'$ORIGIN/../../openairinterface5g_b200' -o $HERE/_build/libnrb200_shim_rx_pusch.so
gcc $F $INC $DEFS $ROOT/oracle/ref_stubs.c $ROOT/oracle/ref_harness_rxpusch.c $HERE/oai_shim_rx_pusch.c $HERE/oai_shim_pusch_chest.c \
    $R/openair1/PHY/NR_ESTIMATION/nr_measurements_gNB.c $R/openair1/PHY/TOOLS/signal_energy.c $R/openair1/PHY/TOOLS/dB_routines.c $R/openair1/PHY/NR_REFSIG/dmrs_nr.c \
    $R/common/utils/nr/nr_common.c $R/openair1/PHY/TOOLS/log2_approx.c $R/openair1/PHY/TOOLS/cmult_sv.c $R/openair1/PHY/NR_TRANSPORT/nr_tbs_tools.c $R/openair1/PHY/NR_REFSIG/ul_ref_seq_nr.c \
    $LIB -Wl,-rpath,'$ORIGIN/../../openairinterface5g_b200' -lm -ldl -Wl,--no-undefined -o $W/libshimtest_rxpusch.so || echo "libshimtest_rxpusch.so: FAILED"
ls -la $HERE/_build/libnrb200_shim_rx_pusch.so $W/libshimtest_rxpusch.so
# the RU front end: nr_feptx0 / nr_fep_full -> the slot-level OFDM entry points of libdfts_b200.so
LIBD="-L$ROOT/openairinterface5g_b200 -l:libdfts_b200.so"
gcc $F $INC $DEFS $HERE/oai_shim_ru_ofdm.c $LIBD -Wl,-rpath,'$ORIGIN/../../openairinterface5g_b200' -o $HERE/_build/libnrb200_shim_ru_ofdm.so
gcc $F $INC $DEFS $ROOT/oracle/ref_harness_ru.c $HERE/oai_shim_ru_ofdm.c $LIBD -Wl,-rpath,'$ORIGIN/../../openairinterface5g_b200' -lm -Wl,--no-undefined \
    -o $W/libshimtest_ru.so || echo "libshimtest_ru.so: FAILED"
ls -la $HERE/_build/libnrb200_shim_ru_ofdm.so $W/libshimtest_ru.so
# the gNB's transport-block decoder: nr_ulsch_decoding.  The reference's definition shares its object file with new_gNB_ulsch / free_gNB_ulsch, which the caller
# keeps using, so the interposer is linked AHEAD of it with --allow-multiple-definition (first definition wins) -- the recipe for a softmodem link line too.
gcc $F $INC $DEFS $HERE/oai_shim_ulsch_decoding.c $LIB -Wl,-rpath,'$ORIGIN/../../openairinterface5g_b200' -o $HERE/_build/libnrb200_shim_ulsch_decoding.so
gcc $F -mpclmul -D_GNU_SOURCE $INC $DEFS $HERE/oai_shim_ulsch_decoding.c $ROOT/oracle/ref_stubs.c $ROOT/oracle/ref_stubs_ulsch.c $ROOT/oracle/ref_harness_ulsch.c \
    $R/openair1/PHY/NR_TRANSPORT/nr_ulsch_decoding.c $R/openair1/PHY/CODING/nr_segmentation.c $R/openair1/PHY/CODING/nr_rate_matching.c $R/openair1/PHY/CODING/crc_byte.c \
    $R/openair1/PHY/NR_TRANSPORT/nr_tbs_tools.c $R/openair1/PHY/TOOLS/dB_routines.c $R/common/utils/threadPool/thread-pool.c \
    -Wl,--allow-multiple-definition $LIB -Wl,-rpath,'$ORIGIN/../../openairinterface5g_b200' -lm -ldl -lpthread -Wl,--no-undefined -o $W/libshimtest_ulsch.so || echo "libshimtest_ulsch.so: FAILED"
ls -la $HERE/_build/libnrb200_shim_ulsch_decoding.so $W/libshimtest_ulsch.so
