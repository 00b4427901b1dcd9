"""Static SASS instruction count per source function for one kernel (from nvdisasm -g output)."""
import re, collections, sys
dis, pattern = sys.argv[1], sys.argv[2]
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if '.section' in l and pattern in l and '.text.' in l)
src = open('/root/repo/coral_b200/csrc/beam_core.h').read().split('\n')
cur = None
cnt = collections.Counter()
for l in lines[start + 1:]:
    if re.match(r'\s*\.section', l):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        cnt[cur] += 1
marks = [(i + 1, l.strip()[:64]) for i, l in enumerate(src) if re.match(r'\s*(static )?CORAL_(DEV|HD|DEV_OUTLINE) ', l)]
marks.append((10 ** 9, 'END'))
fn = collections.Counter()
for (f, ln), c in cnt.items():
    name = f
    if f == 'beam_core.h':
        for (a, nm), (b, _) in zip(marks, marks[1:]):
            if a <= ln < b:
                name = nm
                break
    fn[name] += c
print("total instructions", sum(cnt.values()), "=", sum(cnt.values()) * 16 // 1024, "KB")
for k, v in fn.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 16):
    print(f"{v:6d}  {k}")
