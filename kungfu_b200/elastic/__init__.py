"""kungfu_b200.elastic"""
