from .scheduling_flow_match_euler_discrete import FlowMatchEulerDiscreteScheduler  # noqa: F401
