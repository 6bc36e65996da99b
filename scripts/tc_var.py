import sys, json, subprocess, threading, time
sys.path.insert(0, ".")
from scripts.bench_paths import cfg4
samples = []
stop = False
def poll():
    while not stop:
        o = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active', '--format=csv,noheader'], capture_output=True, text=True).stdout.strip()
        samples.append(o); time.sleep(0.2)
th = threading.Thread(target=poll); th.start()
for C in (148, 148, 64, 148, 296):
    r = cfg4(C=C)
    print(C, round(r['ms'], 2), '%.3e' % r['chain_steps_per_s'], flush=True)
stop = True; th.join()
print(sorted(set(samples))[:12])
