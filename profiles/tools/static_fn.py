# static SASS instruction count per source function for one kernel, from `nvdisasm -g -c` output
import re, sys, bisect, collections
txt = open(sys.argv[1]).read().split('\n')
kern = sys.argv[2]
dev = open('/root/repo/quadrupedal_foothold_planner_b200/csrc/fhp_device.cuh').read().split('\n')
marks = []
for i, l in enumerate(dev, 1):
    if l.startswith('__device__') or l.startswith('template <'):
        nm = re.findall(r'(\w+)\(', l)
        if nm: marks.append((i, nm[0]))
inside = False; cur = ('?', 0); agg = collections.Counter(); lines = collections.Counter(); total = 0
for l in txt:
    if l.startswith('.text.'):
        inside = kern in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,5}\*/', l):
        f, ln = cur
        if f == 'fhp_device.cuh':
            k = bisect.bisect_right([x[0] for x in marks], ln) - 1
            key = marks[k][1] if k >= 0 else 'top'
        else:
            key = f
        agg[key] += 1; lines[(f, ln)] += 1; total += 1
print('total static instructions', total, '=', total * 16 // 1024, 'KB')
for k, v in agg.most_common(30): print(f'{v:6d} {v*16/1024:6.1f} KB  {k}')
print('-- top lines')
for (f, ln), v in lines.most_common(25): print(f'{v:5d} {f}:{ln}')
