class control:
    """Control parameters (mirror of isaacgymenvs/controllers/control_config.py:1-18).

    controller:
        lee_position_control: command_actions = [x, y, z, yaw] in environment frame
        lee_velocity_control: command_actions = [vx, vy, vz, yaw_rate] in vehicle frame
        lee_attitude_control: command_actions = [thrust, roll, pitch, yaw_rate] in vehicle frame
    """
    controller = "lee_position_control"
    kP = [0.8, 0.8, 1.0]
    kV = [0.5, 0.5, 0.4]
    kR = [3.0, 3.0, 1.0]
    kOmega = [0.5, 0.5, 1.20]
    scale_input = [1.0, 1.0, 1.0, 1.0]
