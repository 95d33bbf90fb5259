def get_world_size():
    return 1
