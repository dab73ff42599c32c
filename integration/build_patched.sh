#!/bin/bash
# Build the REFERENCE's own command line with its SGD learner swapped for the libfmb200
# binding (integration/fm_learn_sgd_b200.h).  The reference sources are compiled from
# /root/reference; the two-line edit of libfm.cpp happens on a temporary copy (nothing
# of the reference is stored in this repo).  Output: oracle/_ref/libFM_b200.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF:-/root/reference}"
OUT="$ROOT/oracle/_ref"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT"
sed -e 's|#include "src/fm_learn_sgd_element.h"|#include "src/fm_learn_sgd_element.h"\n#include "fm_learn_sgd_b200.h"|' \
    -e 's|fml = new fm_learn_sgd_element();|fml = new fm_learn_sgd_b200();|' \
    "$REF/src/libfm/libfm.cpp" > "$TMP/libfm_b200.cpp"
grep -q 'fm_learn_sgd_b200()' "$TMP/libfm_b200.cpp"
# MCMC / ALS: the two call sites of the e-term pass (fm_learn_mcmc_simultaneous.h:69,122) go through
# the binding of fmb200_mcmc_eterms when FMB200_MCMC_ETERMS=1, through the reference's own member otherwise
mkdir -p "$TMP/src"
sed -e 's|#include "fm_learn_mcmc.h"|#include "fm_learn_mcmc.h"\n#include "fm_mcmc_eterms_b200.h"|' \
    -e 's|^\( *\)predict_data_and_write_to_eterms(main_data, main_cache);|\1if (b200_eterms_enabled()) b200_predict_data_and_write_to_eterms(fm, main_data, main_cache); else predict_data_and_write_to_eterms(main_data, main_cache);|' \
    "$REF/src/libfm/src/fm_learn_mcmc_simultaneous.h" > "$TMP/src/fm_learn_mcmc_simultaneous.h"
test "$(grep -c 'b200_predict_data_and_write_to_eterms' "$TMP/src/fm_learn_mcmc_simultaneous.h")" = 2
g++ -O3 -w "$TMP/libfm_b200.cpp" -o "$OUT/libFM_b200" \
    -I"$REF/src/libfm" -I"$REF/src/libfm/src" -I"$HERE" -I"$ROOT/include" \
    -L"$ROOT/libfm_b200/lib" -lfmb200 -Wl,-rpath,'$ORIGIN/../../libfm_b200/lib'
echo "$OUT/libFM_b200"
