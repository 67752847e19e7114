"""`pl.plugins.DDPPlugin(find_unused_parameters=False)` / `DDPSpawnPlugin` (scripts/run.py:84-89): markers
— data parallelism is one process per GPU under torchrun with deblur_e_nerf_b200.ddp."""


class DDPPlugin:
    def __init__(self, find_unused_parameters=False, **_kw):
        self.find_unused_parameters = find_unused_parameters


class DDPSpawnPlugin(DDPPlugin):
    pass
