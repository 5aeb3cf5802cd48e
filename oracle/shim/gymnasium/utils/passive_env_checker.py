def check_space(space, space_type, check_box_space_fn=None):
    return None


def check_action_space(space):
    return None


def check_observation_space(space):
    return None
