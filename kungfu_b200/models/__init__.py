"""kungfu_b200.models"""
