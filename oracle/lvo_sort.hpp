/*
 * lvo_sort.hpp -- literal CPU restatement of the PPLL resolve sorts and blend.
 * TEST INFRASTRUCTURE ONLY (see lvo_shaders.hpp header).  PARITY UNPINNED by the reference's own tests.
 *
 * Follows Data/Shaders/Renderers/PPLL/LinkedListSort.glsl:36-263 and LinkedListQuicksort.glsl:29-141.
 * colorList / depthList are the per-invocation arrays of LinkedListResolve.glsl:35-36.
 */
#ifndef LVO_SORT_HPP
#define LVO_SORT_HPP

#include <vector>
#include <algorithm>
#include "lvo_shaders.hpp"

namespace lvo {

struct ResolveLists {
    std::vector<uint32_t> colorList;
    std::vector<float> depthList;
    std::vector<int> stackMemory;
    int stackCounter = 0;

    void swapFragments(uint32_t i, uint32_t j) {  // LinkedListSort.glsl:36-43
        std::swap(colorList[i], colorList[j]);
        std::swap(depthList[i], depthList[j]);
    }

    vec4 blendFTB(uint32_t fragsCount) {  // :45-58
        vec4 color{0, 0, 0, 0};
        for (uint32_t i = 0; i < fragsCount; i++) {
            vec4 colorSrc = unpackUnorm4x8(colorList[i]);
            color.x = color.x + (1.0f - color.w) * colorSrc.w * colorSrc.x;
            color.y = color.y + (1.0f - color.w) * colorSrc.w * colorSrc.y;
            color.z = color.z + (1.0f - color.w) * colorSrc.w * colorSrc.z;
            color.w = color.w + (1.0f - color.w) * colorSrc.w;
        }
        return vec4{color.x / color.w, color.y / color.w, color.z / color.w, color.w};
    }

    vec4 bubbleSort(uint32_t fragsCount) {  // :61-76
        bool changed;
        do {
            changed = false;
            for (uint32_t i = 0; i + 1 < fragsCount; ++i) {
                if (depthList[i] > depthList[i + 1]) { swapFragments(i, i + 1); changed = true; }
            }
        } while (changed);
        return blendFTB(fragsCount);
    }

    void insertionSortOnly(uint32_t fragsCount) {  // :79-102 (without the blend)
        for (uint32_t i = 1; i < fragsCount; ++i) {
            uint32_t fragColor = colorList[i];
            float fragDepth = depthList[i];
            uint32_t j = i;
            while (j >= 1 && depthList[j - 1] > fragDepth) {
                colorList[j] = colorList[j - 1];
                depthList[j] = depthList[j - 1];
                --j;
            }
            colorList[j] = fragColor;
            depthList[j] = fragDepth;
        }
    }
    vec4 insertionSort(uint32_t fragsCount) { insertionSortOnly(fragsCount); return blendFTB(fragsCount); }

    vec4 shellSort(uint32_t fragsCount) {  // :105-136
        const uint32_t gaps[4] = {24, 9, 4, 1};
        for (uint32_t g = 0; g < 4; g++) {
            uint32_t gap = gaps[g];
            for (uint32_t i = gap; i < fragsCount; ++i) {
                uint32_t fragColor = colorList[i];
                float fragDepth = depthList[i];
                uint32_t j = i;
                while (j >= gap && depthList[j - gap] > fragDepth) {
                    colorList[j] = colorList[j - gap];
                    depthList[j] = depthList[j - gap];
                    j -= gap;
                }
                colorList[j] = fragColor;
                depthList[j] = fragDepth;
            }
        }
        return blendFTB(fragsCount);
    }

    void maxHeapSink(uint32_t x, uint32_t fragsCount) {  // :139-156
        uint32_t c;
        while ((c = 2 * x + 1) < fragsCount) {
            if (c + 1 < fragsCount && depthList[c] < depthList[c + 1]) ++c;
            if (depthList[x] >= depthList[c]) return;
            swapFragments(x, c);
            x = c;
        }
    }
    vec4 heapSort(uint32_t fragsCount) {  // :158-171
        for (uint32_t i = (fragsCount + 1) / 2; i > 0; --i) maxHeapSink(i - 1, fragsCount);
        for (uint32_t i = 1; i < fragsCount; ++i) {
            swapFragments(0, fragsCount - i);
            maxHeapSink(0, fragsCount - i);
        }
        return blendFTB(fragsCount);
    }

    void minHeapSink4(uint32_t x, uint32_t fragsCount) {  // :174-199
        uint32_t c, t;
        while ((t = 4 * x + 1) < fragsCount) {
            if (t + 1 < fragsCount && depthList[t] > depthList[t + 1]) c = t + 1; else c = t;
            if (t + 2 < fragsCount && depthList[c] > depthList[t + 2]) c = t + 2;
            if (t + 3 < fragsCount && depthList[c] > depthList[t + 3]) c = t + 3;
            if (depthList[x] <= depthList[c]) return;
            swapFragments(x, c);
            x = c;
        }
    }
    vec4 frontToBackPQ(uint32_t fragsCount) {  // :202-238
        for (uint32_t i = fragsCount / 4; i > 0; --i) minHeapSink4(i, fragsCount);
        vec4 rayColor{0, 0, 0, 0};
        uint32_t i = 0;
        while (i < fragsCount && rayColor.w < 0.99f) {
            minHeapSink4(0, fragsCount - i++);
            vec4 colorSrc = unpackUnorm4x8(colorList[0]);
            rayColor.x = rayColor.x + (1.0f - rayColor.w) * colorSrc.w * colorSrc.x;
            rayColor.y = rayColor.y + (1.0f - rayColor.w) * colorSrc.w * colorSrc.y;
            rayColor.z = rayColor.z + (1.0f - rayColor.w) * colorSrc.w * colorSrc.z;
            rayColor.w = rayColor.w + (1.0f - rayColor.w) * colorSrc.w;
            colorList[0] = colorList[fragsCount - i];
            depthList[0] = depthList[fragsCount - i];
        }
        rayColor.x = rayColor.x / rayColor.w; rayColor.y = rayColor.y / rayColor.w; rayColor.z = rayColor.z / rayColor.w;
        return rayColor;
    }

    // :241-263.  NOTE: as written in the reference this network only sorts when fragsCount is a power of
    // two (the k-loop stops at k <= fragsCount and out-of-range partners are skipped); see DESIGN.md.
    vec4 bitonicSort(uint32_t fragsCount) {
        for (uint32_t k = 2; k <= fragsCount; k *= 2) {
            for (uint32_t j = k / 2; j > 0; j /= 2) {
                for (uint32_t i = 0; i < fragsCount; i++) {
                    uint32_t l = i ^ j;
                    if (l > i && l < fragsCount) {
                        float di = depthList[i], dl = depthList[l];
                        if (((i & k) == 0 && di > dl) || ((i & k) != 0 && di < dl)) swapFragments(i, l);
                    }
                }
            }
        }
        return blendFTB(fragsCount);
    }

    // LinkedListQuicksort.glsl:29-57
    void stackPush(int value) { if (stackCounter < int(stackMemory.size())) { stackMemory[stackCounter] = value; stackCounter++; } }
    int stackPop() { if (stackCounter > 0) { stackCounter--; return stackMemory[stackCounter]; } return 0; }
    bool stackEmpty() const { return stackCounter == 0; }

    int partitionQuicksortLomuto(int low, int high) {  // :59-70
        float pivotElement = depthList[high];
        int i = low;
        for (int j = low; j <= high; j++) {
            if (depthList[j] < pivotElement) { swapFragments(i, j); i++; }
        }
        swapFragments(i, high);
        return i;
    }
    int partitionQuicksortHoare(int low, int high) {  // :72-95
        float e0 = depthList[low], e1 = depthList[(low + high) / 2], e2 = depthList[high];
        float pivotElement = e0 < e1 ? (e2 < e0 ? e0 : fmin_(e1, e2)) : (e2 < e1 ? e1 : fmin_(e0, e2));
        int i = low - 1, j = high + 1;
        while (true) {
            do { i = i + 1; } while (depthList[i] < pivotElement);
            do { j = j - 1; } while (depthList[j] > pivotElement);
            if (i >= j) return j;
            swapFragments(i, j);
        }
    }
    vec4 quicksort(uint32_t fragsCount) {  // :97-118
        stackPush(0); stackPush(int(fragsCount) - 1);
        while (!stackEmpty()) {
            int high = stackPop(); int low = stackPop();
            int pivot = partitionQuicksortLomuto(low, high);
            if (low < pivot - 1) { stackPush(low); stackPush(pivot - 1); }
            if (pivot + 1 < high) { stackPush(pivot + 1); stackPush(high); }
        }
        return blendFTB(fragsCount);
    }
    vec4 quicksortHybrid(uint32_t fragsCount) {  // :120-141
        stackPush(0); stackPush(int(fragsCount) - 1);
        if (fragsCount > 16) {
            while (!stackEmpty()) {
                int high = stackPop(); int low = stackPop();
                int pivot = partitionQuicksortHoare(low, high);
                if (low + 16 < pivot) { stackPush(low); stackPush(pivot - 1); }
                if (pivot + 16 < high) { stackPush(pivot + 1); stackPush(high); }
            }
        }
        insertionSortOnly(fragsCount);
        return blendFTB(fragsCount);
    }

    // sortingAlgorithm dispatch -- src/Renderers/OIT/PerPixelLinkedListLineRenderer.cpp:177-200
    vec4 sortingAlgorithm(int mode, uint32_t fragsCount, uint32_t maxNumFrags) {
        // STACK_SIZE = ceil(log2(MAX_NUM_FRAGS)) * 2 + 4  (PerPixelLinkedListLineRenderer.cpp:173)
        int stackSize = int(std::ceil(std::log2(double(maxNumFrags))) * 2 + 4);
        stackMemory.assign(size_t(stackSize), 0);
        stackCounter = 0;
        switch (mode) {
            case 0: return frontToBackPQ(fragsCount);
            case 1: return bubbleSort(fragsCount);
            case 2: return insertionSort(fragsCount);
            case 3: return shellSort(fragsCount);
            case 4: return heapSort(fragsCount);
            case 5: return bitonicSort(fragsCount);
            case 6: return quicksort(fragsCount);
            default: return quicksortHybrid(fragsCount);
        }
    }

    // Canonical order used by the CUDA product and the parity tests: ascending (depth bits, colour) u64
    // key (depths are non-negative, so the bit pattern orders like the float); blend like blendFTB, or
    // like frontToBackPQ (stop once alpha >= 0.99) for mode 0.  Equals every correct sort above whenever
    // no two fragments of a pixel share a depth.
    vec4 canonical(int mode, uint32_t fragsCount) {
        std::vector<uint64_t> keys(fragsCount);
        for (uint32_t i = 0; i < fragsCount; i++) keys[i] = (uint64_t(f2u(depthList[i])) << 32) | colorList[i];
        std::sort(keys.begin(), keys.end());
        vec4 c{0, 0, 0, 0};
        for (uint32_t i = 0; i < fragsCount; i++) {
            if (mode == 0 && !(c.w < 0.99f)) break;
            vec4 s = unpackUnorm4x8(uint32_t(keys[i] & 0xffffffffu));
            c.x = c.x + (1.0f - c.w) * s.w * s.x;
            c.y = c.y + (1.0f - c.w) * s.w * s.y;
            c.z = c.z + (1.0f - c.w) * s.w * s.z;
            c.w = c.w + (1.0f - c.w) * s.w;
        }
        return vec4{c.x / c.w, c.y / c.w, c.z / c.w, c.w};
    }
};

}  // namespace lvo
#endif
