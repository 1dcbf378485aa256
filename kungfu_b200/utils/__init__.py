"""kungfu_b200.utils"""
