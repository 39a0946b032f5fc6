"""Host front ends under libstdc++ assertions (bounds-checked containers) and UBSan traps, over every golden stream clean,
truncated and corrupted, in both parse modes (device derivations / host derivations).  This image has no libasan / libubsan
/ gdb, so the build traps (-fsanitize-undefined-trap-on-error) and tools/fuzz/trap.h reports the trapping address:
  cd tools/parsebench && g++ -O1 -g -no-pie -include ../fuzz/trap.h -D_GLIBCXX_ASSERTIONS -fsanitize=undefined,bounds \
     -fsanitize-undefined-trap-on-error -std=c++17 -I../../m2dec_b200/csrc -I../../include -o /tmp/pbtrap main.cpp \
     ../../m2dec_b200/csrc/{bits,backend,plugin,validate,mpeg2_front,h264_front}.cpp
  MALLOC_CHECK_=3 MALLOC_PERTURB_=165 python tools/fuzz/parse_fuzz.py <seed> <trials per stream>
  addr2line -f -i -e /tmp/pbtrap <address of a failure>
Last run of round 2: seeds 7 (8 trials) and 11 (20 trials), 1568 runs, no failure -- after two left shifts of negative
values in the MPEG-2 front end (motion vector predictor, intra DC) had been replaced by multiplications."""
import os, random, subprocess, sys, glob
pb='/tmp/pbtrap'
files=sorted(glob.glob('/root/repo/tests/golden/*.264')+glob.glob('/root/repo/tests/golden/*.m2v'))
rng=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 7)
bad=0; runs=0
env=dict(os.environ, ASAN_OPTIONS='detect_leaks=0', UBSAN_OPTIONS='print_stacktrace=1:halt_on_error=1')
for f in files:
    data=open(f,'rb').read()
    codec='mpeg2' if f.endswith('.m2v') else 'h264'
    for trial in range(int(sys.argv[2]) if len(sys.argv)>2 else 12):
        d=bytearray(data)
        if trial==0: pass
        elif trial%3==0: d=d[:rng.randrange(1,len(d))]
        else:
            for _ in range(rng.choice([1,2,8,64])):
                pos=rng.randrange(len(d)); d[pos]=rng.randrange(256) if rng.random()<0.5 else d[pos]^(1<<rng.randrange(8))
        open('/tmp/fz.es','wb').write(bytes(d))
        for mode in ({'B2_PARSE_AS_DEVICE':'1'},{}):
            r=subprocess.run([pb,codec,'/tmp/fz.es','1'],env=dict(env,**mode),stdout=subprocess.PIPE,stderr=subprocess.PIPE,text=True,timeout=30)
            runs+=1
            if r.returncode!=0 or 'runtime error' in r.stderr or 'AddressSanitizer' in r.stderr:
                bad+=1
                print('FAIL',f,trial,mode,r.returncode,r.stderr[-1500:])
                open('/tmp/fz_fail_%d.es'%bad,'wb').write(bytes(d))
                if bad>3: sys.exit(1)
print('runs',runs,'bad',bad)
