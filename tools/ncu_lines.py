"""Per-source-line share of executed warp instructions and stall samples of one kernel in an ncu report
(captured with --set full --import-source on; sources compiled with -lineinfo).
usage: python tools/ncu_lines.py <report.ncu-rep> [kernel regex] [min share]"""
import csv,sys,subprocess,collections
rep=sys.argv[1]; kern=sys.argv[2] if len(sys.argv)>2 else None
cmd=['ncu','-i',rep,'--page','source','--print-source','sass,cuda','--csv']
if kern: cmd+=['--kernel-name','regex:'+kern]
out=subprocess.run(cmd,capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; hdr=None; res=[]
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; continue
    if r[0]=='Function Name' : continue
    if hdr and r[0].isdigit():
        ie=hdr.index('Instructions Executed'); it=hdr.index('Thread Instructions Executed'); ist=hdr.index('Warp Stall Sampling (All Samples)')
        try: res.append((cur,int(r[0]),r[1].strip(),int(r[ie]),int(r[it]),int(r[ist] or 0)))
        except ValueError: pass
tot=sum(x[3] for x in res); tst=sum(x[5] for x in res)
print("total warp instr",tot,"stall samples",tst)
thr=float(sys.argv[3]) if len(sys.argv)>3 else 0.004
for f,l,src,e,t,st in res:
    if e>thr*tot or st>thr*tst:
        print(f"{f[:18]:18s} {l:4d} {e/tot:6.2%} thr/inst={t/max(e,1):5.1f} stall={st/max(tst,1):6.2%}  {src[:110]}")
