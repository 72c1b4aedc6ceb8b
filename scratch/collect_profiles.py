"""gpurun_out/r2f_* (scratch/r2_final.sh on a B200) -> profiles/r02_*: bench lines, ncu summaries, launch lists, probes, sanitizer and
SASS census.  usage: python scratch/collect_profiles.py   (from the repo root, after the record run has been merged back)"""
import json, os, re, subprocess, sys

PARTS = set(sys.argv[1:]) or {'bench', 'ncu', 'ncu_presets', 'launches', 'probes', 'sanitizer', 'sass'}    # which records to refresh
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
G, P, T = 'gpurun_out', 'profiles', 'r2f'
rev = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()


def last_json_line(path):
    for line in reversed(open(path).read().strip().splitlines()):
        line = line.strip()
        if line.startswith('{'):
            return json.loads(line)
    raise ValueError(path)


def sh(cmd, out=None):
    r = subprocess.run(cmd, shell=True, capture_output=True, text=True)
    if out:
        open(out, 'w').write(r.stdout)
    return r.stdout


# ---- bench lines
names = {'bench': 'nerf', 'bench_ref': 'reference_arm', 'bench_nerf_vm': 'nerf_vm', 'bench_nerf_cp': 'nerf_cp', 'bench_image': 'image',
         'bench_sdf': 'sdf', 'bench_image_set': 'image_set', 'bench_nerf_eval': 'nerf_eval'}
for src, dst in (names.items() if 'bench' in PARTS else ()):
    f = f'{G}/{T}_{src}.json'
    try:
        d = last_json_line(f)
        open(f'{P}/r02_bench_{dst}.json', 'w').write(json.dumps(d) + '\n')
        k = d.get('kernels') or {}
        print(f'{dst:14s} {d.get("ms_per_step", 0):8.4f} ms/step  value {d.get("value", 0):.4g} {d.get("unit", "")}  e2e {d.get("e2e", {}).get("value", 0):.4g}  '
              f'roofline {d.get("roofline", {}).get("frac")}  ' + ' '.join(f'{a}={b["ms_per_step"]}' for a, b in k.items()))
    except Exception as e:
        print('MISSING', f, e)

# ---- ncu --set full summaries
rep = f'{G}/{T}_prof_step.ncu-rep'
if 'ncu' not in PARTS:
    pass
elif os.path.exists(rep):
    subprocess.run([sys.executable, 'scratch/ncu_summary.py', rep, f'{P}/r02_ncu_step', f'Round 2: kernels of the nerf.yaml train step (build {rev}), ncu --set full, one launch each',
                    "ncu --set full --clock-control none --import-source on -k regex:'fast_fwd_kernel|fast_bwd_saved_agg|mlp2p_fwd|mlp2p_bwd|rgb_fwd_kernel|rgb_bwd_kernel' -s 36 -c 6 python scratch/prof_step.py"],
                   stdout=subprocess.DEVNULL)
    # the captured launch's query count (bench.py scales the capture's sectors per query by it)
    m = re.search(r'n_valid (\d+) n_app (\d+)', open(f'{G}/{T}_ncu_step.log').read())
    if m:
        d = json.load(open(f'{P}/r02_ncu_step.json'))
        for k, rec in d.items():
            rec['queries_per_launch'] = int(m.group(2) if 'rgb_' in k else m.group(1))
        json.dump(d, open(f'{P}/r02_ncu_step.json', 'w'), indent=1)
    print('ncu summary step')
else:
    print('MISSING', rep)
# the preset / regression captures were summarised on the box (scratch/r2_final.sh: the .ncu-rep files would exceed what gpurun brings back)
for w in (('nerf_vm', 'nerf_cp', 'image', 'sdf', 'image_set') if 'ncu_presets' in PARTS else ()):
    ok = False
    for ext in ('md', 'json'):
        f = f'{G}/{T}_ncusum_{w}.{ext}'
        if os.path.exists(f):
            txt = open(f).read()
            if ext == 'md':
                txt = txt.replace(', ncu --set full, one launch each', f' (build {rev}), ncu --set full, one launch each', 1)
            open(f'{P}/r02_ncu_{w}.{ext}', 'w').write(txt)
            ok = True
    print('ncu summary', w, 'ok' if ok else 'MISSING')

# ---- launch lists
for src, dst, cmd in ((('launches', 'launches', 'python bench.py --steps 2 --warmup 3 --eager'),
                       ('launches_graph', 'launches_graph', 'python bench.py --steps 4 --warmup 3   (the default CUDA-graph step: kernel nodes of the replayed graph)'))
                      if 'launches' in PARTS else ()):
    f = f'{G}/{T}_{src}.csv'
    if os.path.exists(f) and os.path.getsize(f) > 1000:
        body = sh(f'{sys.executable} scratch/launch_summary.py {f}')
        open(f'{P}/r02_{dst}.md', 'w').write(f'# ncu launch list of the nerf.yaml train step (build {rev}): `ncu --metrics gpu__time_duration.sum --clock-control none '
                                             f'-c 200 {cmd}`\n\nPer-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n' + body)
        print('launch list', dst)
    else:
        print('MISSING', f)

# ---- probes
try:
    if 'probes' not in PARTS:
        raise KeyError('skipped')
    open(f'{P}/r02_probe_red.json', 'w').write(json.dumps(last_json_line(f'{G}/{T}_probe_red.json')) + '\n')
    txt = open(f'{G}/{T}_probe_mma.txt').read()
    open(f'{P}/r02_probe_mma.md', 'w').write('# tcgen05.mma issue / execution rate vs N (M = 128, K = 16, bf16, SWIZZLE_NONE smem operands), one elected thread per SM, B200\n\n'
                                             f'`python scratch/probe_mma.py` (ffb_probe_mma, csrc/probe.cu), build {rev}.  The MLPs of this path are made of N = 32 / 64 / 128 MMAs: '
                                             'this table, not the dense peak, is their tensor-pipe roofline (4096 MAC/cycle/SM is reached from N = 128).\n\n```\n' + txt + '```\n')
except Exception as e:
    print('probe', e)

# ---- sanitizer
rows, tails = [], []
for tool in ('memcheck', 'racecheck', 'initcheck'):
    f = f'{G}/{T}_sanitizer_{tool}.log'
    if not os.path.exists(f):
        rows.append(f'| {tool} | missing | |')
        continue
    lines = open(f).read().strip().splitlines()
    rc = next((l.split()[-1] for l in reversed(lines) if l.startswith('sanitizer ')), '?')
    summ = next((l.strip('= ').strip() for l in reversed(lines) if 'SUMMARY' in l), '?')
    rows.append(f'| {tool} | {rc} | {summ} |')
    tails += [f'[{tool}] {l}' for l in lines[-5:]]
if 'sanitizer' in PARTS:
  open(f'{P}/r02_sanitizer.md', 'w').write(
    f'# compute-sanitizer over the shared-memory / vector-reduction / tensor-core kernels (round 2, build {rev})\n\n'
    'Command (per tool): `compute-sanitizer --tool <memcheck|racecheck|initcheck> --error-exitcode 1 python scratch/sanitize_case.py`\n'
    'on one B200.  The script renders 96 rays x 120 samples through `FactorFields.forward` + autograd twice (exact-sized and\n'
    'device-side-count buffers) with the level-parallel dispatch switched off, so the kernels under the tool are the large-batch\n'
    'instantiations the bench times — `fast_fwd_kernel<..,LPAR=0>`, `fast_bwd_saved_agg_kernel` (per-warp run detection, shared-memory\n'
    'parking, vector reductions), the pipelined `mlp2p_fwd/bwd_kernel` (mbarrier rings, tcgen05, TMEM), `rgb_fwd/bwd_kernel`, the\n'
    'composite / sampler kernels — then the -CP preset through `lines_fwd/bwd_kernel` (TMA bulk copies, shared-memory-privatised\n'
    'accumulation), the -vm preset through `vm_fwd2/bwd2_kernel` (warp tiles in shared memory), the 144-channel image preset through\n'
    '`wide_fwd/bwd_kernel`, and the decoupled look-back scan (`scan_tiles_kernel`, 70 000 rows).\n\n'
    '| tool | exit code | summary line |\n|---|---|---|\n' + '\n'.join(rows) + '\n\nLog tails:\n```\n' + '\n'.join(tails) + '\n```\n')

# ---- SASS census (of the library that ran)
if 'sass' in PARTS:
    sh('bash scratch/sass_census.sh', f'{P}/r02_sass_census.md')
print('done; GPU tests:', open(f'{G}/{T}_gpu_tests.log').read().strip().splitlines()[-2:])
