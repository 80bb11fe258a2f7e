// integration/metaeuk_b200_tool.cpp -- a tool built on mmseqs-framework exactly as MetaEuk is (reference: src/metaeuk.cpp:1-120): the framework's base
// commands (createdb, ...) plus the command table entry of the B200 module.  In a MetaEuk checkout the entry below is the one line a maintainer adds to
// `metaeukCommands`; it lives in a table of its own here so that no reference source is copied or modified.
#include "Command.h"
#include "DownloadDatabase.h"
#include "LocalParameters.h"
#include "Prefiltering.h"

const char *binary_name = "metaeuk-b200";
const char *tool_name = "metaeuk-b200";
const char *tool_introduction = "MetaEuk's predictexons hot path on one B200 (libmetaeuk_b200) behind MetaEuk's module surface.";
const char *main_author = "metaeuk_b200";
const char *show_extended_help = "1";
const char *show_bash_info = NULL;
extern const char *MMSEQS_CURRENT_INDEX_VERSION;
const char *index_version_compatible = MMSEQS_CURRENT_INDEX_VERSION;
bool hide_base_commands = true;
bool hide_base_downloads = false;
void (*validatorUpdate)(void) = 0;
std::vector<KmerThreshold> externalThreshold = {};
std::vector<DatabaseDownload> externalDownloads = {};

extern int gpupredictexons(int argc, const char **argv, const Command &command);

LocalParameters &localPar = LocalParameters::getLocalInstance();
std::vector<struct Command> b200Commands = {
        {"gpupredictexons", gpupredictexons, &localPar.predictexonsworkflow, COMMAND_MAIN,
                "Call optimal exon sets based on protein similarity (B200)",
                "predictexons with the six-frame fragments, the k-mer prefilter, the gapped alignment and the exon-chaining DP on the GPU; same DBs in and out.",
                "metaeuk_b200",
                "<i:contigsDB> <i:targetsDB> <o:calledExonsDB>",
                CITATION_METAEUK, {{"contigsDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::nuclDb},
                                   {"targetsDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::sequenceDb},
                                   {"calledExonsDB", DbType::ACCESS_MODE_OUTPUT, DbType::NEED_DATA, NULL}}}
};

extern std::vector<Command> baseCommands;
void init() {
    registerCommands(&baseCommands);
    registerCommands(&b200Commands);
}
void (*initCommands)(void) = init;
void initParameterSingleton() { new LocalParameters; }
