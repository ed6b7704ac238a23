for m in 0 48 40 24; do echo "skip=$m"; PYJAC_DEBUG_SKIP=$m timeout 100 python tools/phase_clocks.py 2>&1 | awk 'NR>2{print $1, $6}' | head -16 | tr '\n' ' '; echo; done
