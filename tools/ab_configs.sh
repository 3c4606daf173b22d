# A/B of library builds on the 16-slot kernel's configs: bash tools/ab_configs.sh lib.so ...
for lib in "$@"; do
  echo "== $lib"
  VIS_B200_LIB=$lib python tools/bench_configs.py 4k_default thumb_4k_2048 thumb_1080p_1024 thumb_4k_1024 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l); print('  ', r['config'], r['images_per_s'], r['hbm_frac'])
    except Exception: print(l.strip()[:300])
"
done
