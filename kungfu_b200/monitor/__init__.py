"""kungfu_b200.monitor"""
