def flatdim(space):  # replaced by gymnasium.spaces.__init__
    raise NotImplementedError
