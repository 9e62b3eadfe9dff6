for c in 3 5 6; do for w in 8 4; do echo "== cap $c world $w"; RTDS_LPT_CAP=$c WORLD=$w ITERS=16 timeout 600 python tools/ab_frame.py lpt=1 2>&1 | cut -c1-230 | tail -1; done; done
RTDS_LPT_CAP=6 ITERS=16 timeout 600 python tools/ab_frame.py lpt=1 2>&1 | cut -c1-230 | tail -1
