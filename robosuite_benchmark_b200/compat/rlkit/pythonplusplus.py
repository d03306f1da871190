from robosuite_benchmark_b200.rlkit_api import list_of_dicts__to__dict_of_lists  # noqa: F401
