for L in "" lib_mb12.so lib_mb16.so; do
  for S in "0.05:1,0.25:2,0.40:4,0.30:8" "0.15:1,0.45:2,0.40:4" "0.25:1,0.50:2,0.25:4" "0.10:1,0.60:2,0.30:4" "0.30:1,0.70:2" "0.05:1,0.55:2,0.40:4" "0.5:1,0.5:2"; do
    echo -n "lib=${L:-default} sched=$S : "; PS_LIB=$L CAPS="$(echo $S | tr ',' ';')" python scripts/he_big.py 2>/dev/null | grep "64x64" | awk '{print $(NF-1), $NF}'
  done
done
