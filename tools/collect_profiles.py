"""Copy the judged artifacts of one tools/round_profile.sh session from gpurun_out/ into profiles/ and write the
per-source-line summaries (the full source-page CSVs are too large to track)."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, out = sys.argv[1], sys.argv[2]            # e.g. r1f r1
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
for src, dst in (('bench_full.json', 'bench_full_2p31rows.json'), ('bench_reference.json', 'bench_reference.json'),
                 ('launches.csv', 'launches.csv'), ('launches_bench.log', 'launches_bench.log'),
                 ('fused_raw.csv', 'fused_raw.csv'), ('cnn_raw.csv', 'cnn_raw.csv'),
                 ('kernel_bench.jsonl', 'kernel_bench.jsonl'), ('pytest.log', 'pytest_gpu.log')):
    shutil.copyfile(os.path.join(G, '%s_%s' % (tag, src)), os.path.join(P, '%s_%s' % (out, dst)))
for src, dst in (('fused_src.csv', 'fused_source_lines.txt'),):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), os.path.join(G, '%s_%s' % (tag, src)), '60'],
                         capture_output=True, text=True).stdout
    with open(os.path.join(P, '%s_%s' % (out, dst)), 'w') as fh:
        fh.write(txt)
print('copied', tag, '->', out)
