"""kungfu_b200.ops"""
