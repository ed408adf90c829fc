"""Same-box A/B: run a tools/ script against several builds of the library, interleaved.
   python tools/ab.py tools/bench_attention.py build/variants/a/libcogaim_b200.so build/variants/b/libcogaim_b200.so"""
import os
import subprocess
import sys
script, libs = sys.argv[1], sys.argv[2:]
for rnd in range(2):
    for lib in libs:
        env = dict(os.environ, CA_LIB_OVERRIDE=os.path.abspath(lib))
        out = subprocess.run([sys.executable, script], env=env, capture_output=True, text=True)
        print(f"== {lib} (round {rnd})\n{out.stdout}{out.stderr[-400:] if out.returncode else ''}", flush=True)
