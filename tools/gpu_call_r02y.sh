set -x
for b in 8 4 2; do for sp in 128 256; do echo "== lpt_bin=$b split=$sp"; RTDS_LPT_BIN=$b RTDS_LPT_SPLIT=$sp WORLD=8 ITERS=14 timeout 600 python tools/ab_frame.py lpt=2 2>&1 | cut -c1-230 | tail -1; done; done
