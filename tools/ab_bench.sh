# same-box A/B of tuning knobs on the bench clip (device-resident leg, e2e, roofline fraction, sequential ms, SM MHz)
run() { python bench.py --no-cpu --no-dropin --steps 16 --warmup 4 "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$TAG', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), round(d['roofline']['sequential_ms_per_step'],3), d['clocks']['sm_mhz'], d['gpu_launches'])"; }
TAG=base run "$@"
TAG=base2 run "$@"
