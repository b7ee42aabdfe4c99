#!/bin/bash
# A/B of the Mean-Teacher loss: two launches (labeled reduce + gradient over everything; gradient kernel generic at git HEAD,
# mode-specialised in the working tree) against the one-launch kernel with its grid barrier; one GPU session.
cd "$(dirname "$0")/.."
run() { echo "== $1"; shift; env "$@" timeout 120 python profiles/loss_microbench.py 2>&1 | tail -2; }
run "HEAD: reduce + generic gradient kernel (2 CTAs/SM)" HPFG_B200_LIB=profiles/ab/libhead.so
run "reduce + specialised gradient kernel (3 CTAs/SM)" HPFG_LOSS_MT_ONE=0 HPFG_LOSS_MT_CTAS_PER_SM=3
run "one launch (labeled sums | unlabeled gradient | barrier | labeled gradient)" HPFG_LOSS_MT_ONE=1
