#!/bin/bash
cd /root/repo
for v in ring3 ring4; do MATE_B200_LIB=/root/repo/scratch/variants/libmate_$v.so python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; done
AB_ARGS=--no-configs bash scratch/ab.sh reg ring3 ring4
