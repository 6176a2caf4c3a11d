"""Tuning sweep: python scripts/sweep.py c3 1024 1024 1024 -- ty=16,ry=2 ty=8,ry=2 ..."""
import os
import subprocess
import sys

args = sys.argv[1:]
sep = args.index('--')
base, tunes = args[:sep], args[sep + 1:]
here = os.path.dirname(os.path.abspath(__file__))
for t in tunes:
    env = dict(os.environ, PSAD_TUNE='' if t == 'default' else t)
    out = subprocess.run([sys.executable, os.path.join(here, 'kbench.py')] + base, env=env, capture_output=True, text=True)
    for line in (out.stdout + out.stderr).splitlines():
        if 'march' in line or 'Error' in line or 'error' in line:
            print('%-40s %s' % (t, line), flush=True)
