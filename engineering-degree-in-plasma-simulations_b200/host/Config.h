// Facade of ch4/v3/src/Config.h: the same singleton flags.  MULTITHREADING / NUM_THREADS are accepted and stored
// but have no effect: the particle loop runs on the GPU.
#ifndef CONFIG_H
#define CONFIG_H
#include <cstddef>
#include <vector>

class Config {
public:
    static Config& getInstance();
    bool setSUBCYCLING(bool o) { return m_SUBCYCLING = o; }
    bool getSUBCYCLING() const { return m_SUBCYCLING; }
    bool setMULTITHREADING(bool o) { return m_MULTITHREADING = o; }
    bool getMULTITHREADING() const { return m_MULTITHREADING; }
    unsigned int setNUM_THREADS(unsigned int n) { return m_NUM_THREADS = n; }
    unsigned int getNUM_THREADS() const { return m_NUM_THREADS; }
    bool setMERGING(bool o) { return m_MERGING = o; }
    bool getMERGING() const { return m_MERGING; }
    bool setINFLUENCE_ON_BACKGROUND(bool o) { return m_INFLUENCE_ON_BACKGROUND = o; }
    bool getINFLUENCE_ON_BACKGROUND() const { return m_INFLUENCE_ON_BACKGROUND; }
    bool setSPUTTERING(bool o) { return m_SPUTTERING = o; }
    bool getSPUTTERING() const { return m_SPUTTERING; }
    Config(const Config&) = delete;
    void operator=(const Config&) = delete;

private:
    Config();
    bool m_SUBCYCLING = true, m_MULTITHREADING = true, m_MERGING = true, m_INFLUENCE_ON_BACKGROUND = false, m_SPUTTERING = false;
    unsigned int m_NUM_THREADS = 1;
};
std::vector<size_t> splitIntoChunks(size_t size);
#endif
