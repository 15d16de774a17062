"""Copy the judged artifacts of one tools/round_profile.sh session (gpurun_out/<tag>_*) into profiles/<out>_*."""
import glob
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, out = sys.argv[1], sys.argv[2]            # e.g. r2a r2
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
keep = ('_raw.csv', '_source_lines.txt', '.json', '.jsonl', '_launches.csv', '_launches_bench.log', '_pytest_gpu.log')
for src in sorted(glob.glob(os.path.join(G, tag + '_*'))):
    name = os.path.basename(src)[len(tag):]
    if name.endswith(keep) and os.path.getsize(src) > 0:
        shutil.copyfile(src, os.path.join(P, out + name))
        print('copied', os.path.basename(src), '->', out + name)
