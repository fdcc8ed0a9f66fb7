"""x500 single-rigid-body constants derived from the reference's assets/x500/x500.urdf.

  base link   mass 2.0, I = diag(0.0216667, 0.0216667, 0.04) at the link origin   (urdf:32-36)
  4 rotors    mass 0.0160769, I = diag(3.846e-7, 2.6116e-5, 2.6499e-5)             (urdf:99-103)
  rotor joints at (0.174,-0.174,0.3) (-0.174,0.174,0.3) (0.174,0.174,0.3) (-0.174,-0.174,0.3)  (urdf:6,13,20,27)

The rotors are free-spinning z-revolute joints (isaacgymenvs/tasks/ouzelum.py:160-163) spun in +/- pairs,
so the kernel integrates ONE rigid body: composite mass, centre of mass offset along body z, composite
inertia about the centre of mass with the blade Ixx/Iyy averaged over a revolution.
The same numbers are compiled into `ozl_cfg_default` (csrc/quad_step.cu); tests check they agree.
"""
import math

M_BASE = 2.0
I_BASE = (0.02166666666666667, 0.02166666666666667, 0.04000000000000001)
M_ROTOR = 0.016076923076923075
I_ROTOR = (3.8464910483993325e-07, 2.6115851691700804e-05, 2.649858234714004e-05)
ROTOR_XY = ((0.174, -0.174), (-0.174, 0.174), (0.174, 0.174), (-0.174, -0.174))
ROTOR_Z = 0.3
ARM = 0.174

MASS = M_BASE + 4.0 * M_ROTOR
COM_Z = 4.0 * M_ROTOR * ROTOR_Z / MASS
_DZ = ROTOR_Z - COM_Z
_IRXY = 0.5 * (I_ROTOR[0] + I_ROTOR[1])
IXX = I_BASE[0] + M_BASE * COM_Z ** 2 + 4.0 * (_IRXY + M_ROTOR * (ARM * ARM + _DZ * _DZ))
IYY = I_BASE[1] + M_BASE * COM_Z ** 2 + 4.0 * (_IRXY + M_ROTOR * (ARM * ARM + _DZ * _DZ))
IZZ = I_BASE[2] + 4.0 * (I_ROTOR[2] + M_ROTOR * (2.0 * ARM * ARM))
MAX_ANGVEL = 4.0 * math.pi      # isaacgymenvs/tasks/ouzelum.py:141
GRAVITY_Z = -9.81               # isaacgymenvs/tasks/ouzelum.py:118
