for f in rl_rocket_6dof_b200/lib/libr6dof_t*.so; do
  R6_LIB_PATH=$f python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('$f', 'step %.3f ms  %.3e/s | rollout %.3f ms | e2e %.3f ms' % (d['ms_per_step'], d['value'], d['rollout_fused']['ms_per_step'], d['e2e']['ms_per_step']))"
done
