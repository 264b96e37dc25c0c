"""Plain-Python equivalents of the reference's ROS messages (src/dql_multirotor_landing/msg/Action.msg,
Observation.msg): same field names, float64 semantics."""


class Observation:
    __slots__ = ("rel_p_x", "rel_p_y", "rel_p_z", "rel_v_x", "rel_v_y", "rel_v_z", "rel_a_x", "rel_a_y", "rel_a_z", "contact")

    def __init__(self, rel_p_x=0.0, rel_p_y=0.0, rel_p_z=0.0, rel_v_x=0.0, rel_v_y=0.0, rel_v_z=0.0,
                 rel_a_x=0.0, rel_a_y=0.0, rel_a_z=0.0, contact=False):
        self.rel_p_x, self.rel_p_y, self.rel_p_z = rel_p_x, rel_p_y, rel_p_z
        self.rel_v_x, self.rel_v_y, self.rel_v_z = rel_v_x, rel_v_y, rel_v_z
        self.rel_a_x, self.rel_a_y, self.rel_a_z = rel_a_x, rel_a_y, rel_a_z
        self.contact = contact


class Action:
    __slots__ = ("roll", "pitch", "yaw", "v_z")

    def __init__(self, roll=0.0, pitch=0.0, yaw=0.0, v_z=0.0):
        self.roll, self.pitch, self.yaw, self.v_z = roll, pitch, yaw, v_z
