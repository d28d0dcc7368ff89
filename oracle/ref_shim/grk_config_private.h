/* see grk_config.h in this directory */
#pragma once
#define GROK_PLUGIN_NAME "grokj2k_plugin"
