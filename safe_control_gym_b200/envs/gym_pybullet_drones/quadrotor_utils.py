'''Quadrotor helpers (safe_control_gym/envs/gym_pybullet_drones/quadrotor_utils.py).'''

from enum import IntEnum


class QuadType(IntEnum):
    ONE_D = 1     # motion along z
    TWO_D = 2     # motion in the x-z plane
    THREE_D = 3


# Crazyflie 2.x constants of assets/cf2x.urdf:5,11-12,42-78 (parsed by base_aviary.py:542-580 in the reference)
CF2X = dict(
    arm=0.0397, kf=3.16e-10, km=7.94e-12, thrust2weight=2.25, mass=0.027,
    ixx=1.4e-5, iyy=1.4e-5, izz=2.17e-5,
    pwm2rpm_scale=0.2685, pwm2rpm_const=4070.3, pwm_min=20000.0, pwm_max=65535.0,
    prop_offset=0.028,                  # |x| = |y| of the prop links' inertial origins
    collision_h=0.025, collision_r=0.06, collision_z_offset=0.0,
    # aerodynamic coefficients of the non-default physics variants (cf2x.urdf:5; base_aviary.py:386-460)
    gnd_eff_coeff=11.36859, prop_radius=2.31348e-2, drag_coeff_xy=9.1785e-7, drag_coeff_z=10.311e-7,
    dw_coeff_1=2267.18, dw_coeff_2=0.16, dw_coeff_3=-0.11,
)
