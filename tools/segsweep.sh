for seg in 12 24 48 96 4000; do
  f=$((seg*65536))
  echo "== seg $seg"
  python tools/gemm_bench.py --layers dec0.c7d1 dec1.c7d1 dec2.c7d1 dec3.c7d1 dec3.c1 dec2.convt --windows 4 --flags $f 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['layer'], d['ms'], d['tflops'])
"
  VOC_TC_FLAGS=$f python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "full_chunk" 2>&1 | grep -E "full/64|passed|failed"
done
