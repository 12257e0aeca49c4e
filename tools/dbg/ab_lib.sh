for i in 1 2 3; do
  for v in old b200; do
    IIC_LIB=$PWD/ai-interior-image-classifier_b200/_lib/libiic_$v.so python bench.py --legs "" --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(v,4) for k,v in d['roofline']['gemm_ms_per_launch_by_shape'].items()})"
  done
done
