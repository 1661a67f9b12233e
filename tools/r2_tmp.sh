timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_reference_tests.py tests/test_gpu_integrators.py tests/test_gpu_scale.py -m gpu -x -q -k "sol or Sol or 256k or golden" 2>&1 | tail -4
python tools/config_timings.py --steps 400 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%-40s N %8d  %.4g atom-steps/s  ms/step %.5f  k1 %.4f force %.4f k3 %.4f rebuild %.3f x%d' % (d['config'][:40], d['n_atoms'], d['atom_steps_per_s'], d['ms_per_step'], d['k1_ms'], d['force_ms'], d['k3_ms'], d['rebuild_ms_each'], d['rebuilds']))
"
