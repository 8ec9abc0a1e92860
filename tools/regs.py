import re,sys
log=open(sys.argv[1]).read()
pat=re.compile(r"Compiling entry function '_ZN4fheb\d+(\w+?)ILi(\d+)ELb(\d)ELi(\d+)ELi(\d+)(?:ELb(\d))?E[^']*' for 'sm_100a'\n.*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*Used (\d+) registers")
for m in pat.finditer(log):
    k,L,lazy,thr,ppc,flag,stack,ss,sl,regs=m.groups()
    if len(sys.argv)<3 or int(L) in map(int,sys.argv[2].split(',')):
        print(f"{k:22s} L={L:>2} lazy={lazy} thr={thr:>4} ppc={ppc:>3} flag={flag} regs={regs:>3} stack={stack} spill={ss}/{sl}")
