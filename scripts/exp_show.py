import json, sys
for d in sys.argv[1:]:
    try:
        j = json.loads(open(f'gpurun_out/x_{d}.json').read().strip().splitlines()[-1])
        k = j['roofline']['kernels']
        print(d, round(j['ms_per_step'], 3), {n: round(v['ms_per_step'], 3) for n, v in k.items()})
    except Exception as e:
        print(d, 'ERR', e)
