for S in "1.0:32" "1.0:24" "1.0:16" "0.05:8,0.95:32" "0.05:16,0.95:32" "0.2:16,0.8:32"; do echo "SCHED=$S"; PWB_HE_SCHED=$S python scripts/time_stages_cfg3.py | tail -1; done
for C in 16 64; do echo "COARSE=$C"; PWB_HE_COARSE=$C python scripts/time_stages_cfg3.py | tail -1; done
