from .plugin import (TRT_LLM_PLUGIN_NAMESPACE, PluginConfig, TrtPlugin, _load_plugin_lib,  # noqa: F401
                     get_plugin_creator)
