for lib in default mttold default mttold; do
  if [ $lib = default ]; then unset PZ_LIB_PATH; else export PZ_LIB_PATH=$PWD/build/variants/libpz_$lib.so; fi
  python bench.py --model mtt --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][0]); print('$lib', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
