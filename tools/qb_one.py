import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo/tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32)
q.run('cartpole', 4096, 200, 64, 'f64', 32)
