#!/bin/bash
# TEST INFRASTRUCTURE. Quirk Q4: the reference reads m_filter_beg[num_banks + 2], one int past the array
# (mfcccpu.cpp:206 against :28). Builds the reference's CPU sources in place with AddressSanitizer (recover mode) and runs
# one golden case: ASan must report exactly that read, and the output must equal the plain build's bit for bit.
# Needs /root/reference; run from the repo root:  bash oracle/asan_check.sh
set -e
REF=${REF:-/root/reference}; OUT=${OUT:-/tmp/afe_asan}; mkdir -p $OUT
gcc -O1 -g -fPIC -ffp-contract=off -fsanitize=address -fsanitize-recover=address -c -o $OUT/shim.o oracle/fftw_shim.c
g++ -std=c++14 -O1 -g -fPIC -w -ffp-contract=off -include cfloat -include stdlib.h -include math.h -I$REF/include -I$REF \
    -fsanitize=address -fsanitize-recover=address -shared -o $OUT/libref_asan.so oracle/ref_capi.cpp \
    $REF/parambase.cpp $REF/mfccbase.cpp $REF/segmentercpu.cpp $REF/deltacpu.cpp $REF/normalizercpu.cpp $REF/mfcccpu.cpp $OUT/shim.o -lpthread
cat > $OUT/run.py <<PY
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle_lib as ol
from golden_io import load_pcm
for v in vars(ol).values():
    if isinstance(v, dict) and isinstance(v.get("ref"), str) and v["ref"].endswith("libref_mfcc.so"):
        v["ref_asan"] = "$OUT/libref_asan.so"
p = ol.default_params(norm="cmn", dyn="acc")
pcm = load_pcm()["a1"]
a = ol.RefLib("ref").extract(p, [pcm], sample_limit=1 << 22)[0][0]
b = ol.RefLib("ref_asan").extract(p, [pcm], sample_limit=1 << 22)[0][0]
print("asan build == plain build:", np.array_equal(a, b), a.shape)
PY
ASAN_OPTIONS=halt_on_error=0:detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) python $OUT/run.py 2>&1 | grep "asan build\|SUMMARY" | sort | uniq -c
