"""Hot source lines of an `ncu --page source --csv --print-source cuda,sass` dump (stall samples per line).
usage: python scripts/ncu_source_hot.py dump.csv [block]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
blocks = [i for i, r in enumerate(rows) if r and r[0] == 'File Path']
want = int(sys.argv[2]) if len(sys.argv) > 2 else None
for b, (s, e) in enumerate(zip(blocks, blocks[1:] + [len(rows)])):
    if want is not None and b != want:
        continue
    hdr = rows[s + 2]
    n = len(hdr)
    neg = {h: j - n for j, h in enumerate(hdr)}  # index from the end: source text may split columns
    def g(r, h):
        try:
            return int(r[neg[h]])
        except Exception:
            return 0
    lines = [r for r in rows[s + 3:e] if r and r[0].isdigit() and len(r) >= n]
    tot = sum(g(r, '# Samples') for r in lines)
    print(f"==== block {b}: {rows[s][1].split('/')[-1]} :: {rows[s+1][1][:50]}  samples={tot} long_sb={sum(g(r,'stall_long_sb') for r in lines)} "
          f"short_sb={sum(g(r,'stall_short_sb') for r in lines)} wait={sum(g(r,'stall_wait') for r in lines)} inst={sum(g(r,'Instructions Executed') for r in lines)}")
    lines.sort(key=lambda r: -g(r, '# Samples'))
    for r in lines[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
        inst = g(r, 'Instructions Executed'); thr = g(r, 'Thread Instructions Executed')
        print(f"{r[0]:>5} smp={g(r,'# Samples'):6d} long={g(r,'stall_long_sb'):6d} short={g(r,'stall_short_sb'):5d} wait={g(r,'stall_wait'):5d} "
              f"br={g(r,'stall_branch_resolving'):4d} thr/inst={thr/max(inst,1):5.1f} inst={inst:9d} | {' '.join(r[1:len(r)-n+2])[:90]}")
