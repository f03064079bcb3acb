from drivescenegen_b200.hostapi.optimization import get_cosine_schedule_with_warmup  # noqa: F401
