import re,sys
lines=[l for l in open(sys.argv[1]) if re.match(r'\s+/\*[0-9a-f]{4}\*/', l)]
ins=[re.sub(r'/\*.*?\*/','',l).strip().rstrip(';').strip() for l in lines]
def regs(tok):
    return [int(x) for x in re.findall(r'\bR(\d+)\b', tok)]
for i,s in enumerate(ins):
    m=re.search(r'LDG\.E\.128\S*\s+R(\d+),', s)
    if not m: continue
    base=int(m.group(1)); dst=set(range(base,base+4))
    for j in range(i+1,min(i+60,len(ins))):
        t=ins[j]
        if t.startswith('@'): t=t.split(None,1)[1]
        parts=t.split(None,1)
        if len(parts)<2: continue
        ops=parts[1].split(',')
        d=regs(ops[0]); srcs=[r for o in ops[1:] for r in regs(o)]
        if 'LDG' in parts[0] or 'BSYNC' in parts[0] or 'BSSY' in parts[0] or 'BRA' in parts[0]: 
            if 'LDG' in parts[0]: continue
            continue
        if set(srcs)&dst: break            # first real use
        if d and d[0] in dst and not parts[0].startswith('ST'):
            print(f"WAW hazard: [{i}] {s}   <-   [{j}] {ins[j]}"); break
