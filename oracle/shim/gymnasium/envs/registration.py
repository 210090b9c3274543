class EnvSpec:
    def __init__(self, id, entry_point=None, **kwargs):
        self.id = id
        self.entry_point = entry_point
        self.kwargs = kwargs
        self.max_episode_steps = kwargs.get("max_episode_steps")
