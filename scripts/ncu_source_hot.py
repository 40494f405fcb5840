"""Top stall sites of an `ncu --page source --csv` dump: python scripts/ncu_source_hot.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix['# Samples']]) for r in data)
texec = sum(int(r[ix['Instructions Executed']]) for r in data)
print('total samples', tot, 'warp instructions', texec)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(range(len(data)), key=lambda k: -int(data[k][ix['# Samples']]))[:N]
for k in sorted(top):
    r = data[k]
    st = {s[6:]: int(r[ix[s]]) for s in stalls if int(r[ix[s]]) > 0.15 * max(1, int(r[ix['# Samples']]))}
    print(k, r[ix['Source']].strip()[:70].ljust(70), 'samples', r[ix['# Samples']], 'exec', r[ix['Instructions Executed']], st)
