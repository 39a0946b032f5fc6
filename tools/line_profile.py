"""Per-source-line view of an `ncu --set full --import-source on` capture (the CLI's source page has no CUDA-line
metrics in CSV form): joins the SASS page of the report with the line table of the cubin.
    cuobjdump -xelf all m2dec_b200/libm2dec_b200.so && nvdisasm -g -c h264_recon.sm_100a.cubin > h264.sass
    ncu -i report.ncu-rep --page source --csv [--kernel-name regex:...] > kernel.csv
    python tools/line_profile.py h264.sass kernel.csv <function name pattern> m2dec_b200/csrc/h264_recon.cu
The library must be the build that was profiled."""
import csv, re, sys, collections
sass_file, csv_file, func_pat, src_file = sys.argv[1:5]
# 1. nvdisasm -g: instruction index -> source line (innermost inlined-at chain ignored: use the line marker)
lines=open(sass_file).read().split('\n')
in_f=False; cur=None; order=[]
for l in lines:
    if l.startswith('.text.'):
        in_f = re.search(func_pat, l) is not None
        continue
    if not in_f: continue
    m=re.search(r'//## File ".*?", line (\d+)', l)
    if m: cur=int(m.group(1)); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: order.append((int(m.group(1),16), cur, m.group(2)))
# 2. ncu csv
rows=list(csv.reader(open(csv_file)))
H=rows[1]
ci=H.index('Instructions Executed'); si=H.index('Source'); smp=H.index('# Samples'); ai=H.index('Address')
stall_cols=[i for i,h in enumerate(H) if h.startswith('stall_')]
base=None; data=[]
for r in rows[2:]:
    try: n=int(r[ci])
    except: continue
    a=int(r[ai],16)
    if base is None: base=a
    data.append((a-base, n, int(r[smp] or 0), r[si], {H[i]:int(r[i] or 0) for i in stall_cols}))
print('sass instrs', len(order), 'csv instrs', len(data))
off2line={o:(ln,txt) for o,ln,txt in order}
src=open(src_file).read().split('\n')
by_line=collections.defaultdict(lambda:[0,0,collections.Counter()])
tot_i=tot_s=0
for off,n,s,txt,st in data:
    ln=off2line.get(off,(None,''))[0]
    by_line[ln][0]+=n; by_line[ln][1]+=s; by_line[ln][2].update(st)
    tot_i+=n; tot_s+=s
print('total instr',tot_i,'samples',tot_s)
print('== top lines by samples')
for ln,(n,s,st) in sorted(by_line.items(), key=lambda kv:-kv[1][1])[:60]:
    top=', '.join('%s %d'%(k[6:],v) for k,v in st.most_common(3) if v)
    print('%5s  inst %5.1f%%  smp %5.1f%%  [%s]  %s'%(ln,100*n/tot_i,100*s/tot_s,top,(src[ln-1].strip()[:90] if ln else '')))
# by region
def region(ln):
    return ln
