"""kungfu_b200.parallel"""
