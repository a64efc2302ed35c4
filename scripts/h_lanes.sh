for L in 1 2 3 4 5 7 8 16; do echo "PWB_H_LANES=$L"; PWB_H_LANES=$L python scripts/time_stages.py 64 2>&1 | grep "rep 2"; done
