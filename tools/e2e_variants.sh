# e2e timeline of pz_inflate_batch_contig with host buffers, for a few flag combinations (PZ_TRACE)
for f in ${E2E_FLAGS:-0 8}; do
  echo "== flags $f"
  PZ_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --verify 0 --e2e-flags $f 2> gpurun_out/e2e_$f.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['e2e']['value'])"
  tail -${E2E_TAIL:-70} gpurun_out/e2e_$f.err
done
