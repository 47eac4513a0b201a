for v in "" _c16 _c8 _c4 ""; do echo "variant $v"; export PTB200_LIB=$PWD/ascendpathtracing_b200/libptb200$v.so
python bench.py --steps 30 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['pipeline']['ms_per_step'])"
python tools/suite_multi_gpu.py --only c4,c5,c5mat --scale 0.25 --out gpurun_out/s.json 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['seconds'],4), round(d['mpaths_s']), round(d['grays_s'],2))"
done
