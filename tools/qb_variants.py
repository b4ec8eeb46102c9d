'''Developer timing of rollout-kernel variants: qb_variants.py [E:NT ...] (GPU box)'''
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import quick_bench as q

variants = [a for a in sys.argv[1:] if ':' in a] or ['32:0', '16:0', '16:256', '8:128']
names = [a for a in sys.argv[1:] if ':' not in a] or ['quadrotor_2D_track']
for name in names:
    for N in ((8192, 65536) if '3D' in name else (4096, 32768)):
        for v in variants:
            E, NT = (int(x) for x in v.split(':'))
            for prec in ('f64', 'f32'):
                q.run(name, N, 200, 64, prec, E | (NT << 16))
