"""kungfu_b200.cmd"""
