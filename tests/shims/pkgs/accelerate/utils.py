class ProjectConfiguration:
    """accelerate.utils.ProjectConfiguration: a plain record of project_dir / logging_dir (the trainer only stores it)."""
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)
