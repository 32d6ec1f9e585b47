"""Rebuilds profiles/traffic.json from the committed ncu raw pages: DRAM bytes per launch (read + write) of the kernels
bench.py reports a roofline for, each tied to the sha of the source file the capture was taken from (bench.py emits
`traffic: null` when the source has changed since).  Run after tools/gpu_final2.sh with the captures copied to profiles/."""
import csv, hashlib, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def sha16(path):
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()[:16]


def launches(fn):
    rows = list(csv.reader(open(os.path.join(ROOT, 'profiles', fn))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        yield dict(name=d['Kernel Name'], grid=d.get('launch__grid_size'),
                   bytes=int(float(d['dram__bytes_read.sum']) * UNIT[u['dram__bytes_read.sum']]
                             + float(d['dram__bytes_write.sum']) * UNIT[u['dram__bytes_write.sum']]),
                   ms=float(d['gpu__time_duration.sum']) * {'us': 1e-3, 'ms': 1.0, 'ns': 1e-6, 's': 1e3}[u['gpu__time_duration.sum']])


def pick(fn, prefix, by='ms'):
    c = [l for l in launches(fn) if l['name'].startswith(prefix) or ('void ' + prefix) in l['name']]
    return max(c, key=lambda l: l[by]) if c else None


out = {'_comment': 'dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full --clock-control none` captures '
                   '(profiles/r02_final_*_ncu_raw.csv); an entry is used by bench.py only while `source` still has the sha16 it '
                   'had when the capture was taken (tools/update_traffic.py)'}
for key, fn, prefix, src in (('mt_tc_interact_kernel', 'r02_final_cfg2_step_ncu_raw.csv', 'mt_tc_interact_kernel', 'mt_tc.cu'),
                             ('lstm_tc_kernel', 'r02_final_cfg2_step_ncu_raw.csv', 'lstm_tc_kernel', 'lstm_tc.cu'),
                             ('rnn_tc_kernel', 'r02_final_cars_gemm_rnn_ncu_raw.csv', 'rnn_tc_kernel', 'rnn_tc.cu'),
                             ('gemm_tc_kernel', 'r02_final_cars_gemm_rnn_ncu_raw.csv', 'gemm_tc2_kernel', 'gemm_tc.cu'),
                             ('drmm_tc_kernel', 'r02_final_drmm_tc_ncu_raw.csv', 'drmm_tc_kernel', 'drmm_tc.cu')):
    # the GEMM entry is the CARS document pre-gate GEMM = the launch that moves the most DRAM bytes (the attention row-dot GEMM of
    # the same capture runs longer but stores nothing)
    l = pick(fn, prefix, by='bytes' if key == 'gemm_tc_kernel' else 'ms')
    if l is None:
        continue
    path = os.path.join('context_attentive_ir_b200', 'csrc', src)
    out[key] = dict(bytes=l['bytes'], ms=l['ms'], capture='profiles/%s (%s, grid %s: the longest / for the GEMM the heaviest launch of that kernel in the capture)' % (fn, l['name'][:40], l['grid']),
                    source=path, source_sha16=sha16(os.path.join(ROOT, path)))
json.dump(out, open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w'), indent=1)
print(json.dumps(out, indent=1))
