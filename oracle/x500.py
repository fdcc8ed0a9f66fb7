"""x500 single-rigid-body constants -- TEST INFRASTRUCTURE (oracle twin of ouzelum_b200/x500.py).

Derived (float64) from the numbers in the reference's assets/x500/x500.urdf:
  base link   mass 2.0, I = diag(0.0216667, 0.0216667, 0.04) at the link origin   (urdf:32-36)
  4 rotors    mass 0.0160769, I = diag(3.846e-7, 2.6116e-5, 2.6499e-5)             (urdf:99-103)
  rotor joints at (0.174,-0.174,0.3) (-0.174,0.174,0.3) (0.174,0.174,0.3) (-0.174,-0.174,0.3)
                                                                                (urdf:6,13,20,27)
The rotors sit on free-spinning z-revolute joints (isaacgymenvs/tasks/ouzelum.py:160-163), spin
in +/- pairs (net angular momentum 0), so the vehicle is integrated as ONE rigid body: composite
mass, COM offset along body z, composite inertia about the COM (rotor blade Ixx/Iyy averaged over
a revolution).  See SURVEY.md section 8a row P.
"""
import math

M_BASE = 2.0
I_BASE = (0.02166666666666667, 0.02166666666666667, 0.04000000000000001)
M_ROTOR = 0.016076923076923075
I_ROTOR = (3.8464910483993325e-07, 2.6115851691700804e-05, 2.649858234714004e-05)
ROTOR_XY = ((0.174, -0.174), (-0.174, 0.174), (0.174, 0.174), (-0.174, -0.174))
ROTOR_Z = 0.3
ARM = 0.174
# propeller handedness from the urdf visuals (1345_prop_ccw for rotors 0,1; _cw for rotors 2,3)
ROTOR_SPIN = (1.0, 1.0, -1.0, -1.0)

MASS = M_BASE + 4.0 * M_ROTOR
COM_Z = 4.0 * M_ROTOR * ROTOR_Z / MASS


def _composite_inertia():
    i_r_xy = 0.5 * (I_ROTOR[0] + I_ROTOR[1])
    ixx = I_BASE[0] + M_BASE * COM_Z ** 2
    iyy = I_BASE[1] + M_BASE * COM_Z ** 2
    izz = I_BASE[2]
    for (x, y) in ROTOR_XY:
        dz = ROTOR_Z - COM_Z
        ixx += i_r_xy + M_ROTOR * (y * y + dz * dz)
        iyy += i_r_xy + M_ROTOR * (x * x + dz * dz)
        izz += I_ROTOR[2] + M_ROTOR * (x * x + y * y)
    return ixx, iyy, izz


IXX, IYY, IZZ = _composite_inertia()
MAX_ANGVEL = 4.0 * math.pi          # isaacgymenvs/tasks/ouzelum.py:141
GRAVITY_Z = -9.81                   # isaacgymenvs/tasks/ouzelum.py:118 / cfg/task/Ouzelum.yaml:23
